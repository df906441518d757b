"""CPU restatement of the reference replay buffers.  TEST INFRASTRUCTURE ONLY.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline /
`--impl reference` leg may import this; the product (`dopamine_b200`) never does.

Restates (file:line relative to /root/reference/dopamine/replay_memory/):
  * invalid cursor window ............. circular_replay_buffer.py:53-77
  * storage / discount vector ......... circular_replay_buffer.py:98-192
  * add + episode-start zero padding .. circular_replay_buffer.py:225-287
  * validity of a transition .......... circular_replay_buffer.py:381-414
  * uniform index sampling ............ circular_replay_buffer.py:436-477
  * batch assembly (stacks, n-step) ... circular_replay_buffer.py:479-558
  * prioritized add / sampling / set .. prioritized_replay_buffer.py:117-235

Parity is PINNED by `tests/test_oracle_golden.py`: the reference's own
known-answer tests (tests/dopamine/replay_memory/*_test.py) re-run against this
port, plus fixtures made by executing the unmodified reference
(`oracle/make_golden.py`).

The per-element Python loops are kept on purpose: this port is also the CPU
baseline timed by `bench.py`, and the reference's cost is dominated by exactly
those loops.
"""
import collections
import math

import numpy as np

from oracle.sumtree_port import PortSumTree

Element = collections.namedtuple('Element', ['name', 'shape', 'type'])


def cursor_window(cursor, capacity, stack_size, horizon):
  """Indices that straddle the write cursor (circular_replay_buffer.py:53-77)."""
  assert cursor < capacity
  first = cursor - horizon
  return np.array([(first + k) % capacity
                   for k in range(stack_size + horizon)])


class PortReplay(object):
  """Uniform circular replay (port of OutOfGraphReplayBuffer)."""

  def __init__(self, observation_shape, stack_size, replay_capacity, batch_size,
               update_horizon=1, gamma=0.99, max_sample_attempts=1000,
               extra_storage_types=None, observation_dtype=np.uint8,
               terminal_dtype=np.uint8, action_shape=(), action_dtype=np.int32,
               reward_shape=(), reward_dtype=np.float32, np_rng=None):
    assert isinstance(observation_shape, tuple)
    if replay_capacity < update_horizon + stack_size:
      raise ValueError('There is not enough capacity to cover '
                       'update_horizon and stack_size.')
    self.obs_shape = observation_shape
    self.stack = stack_size
    self.capacity = replay_capacity
    self.batch = batch_size
    self.horizon = update_horizon
    self.gamma = gamma
    self.max_attempts = max_sample_attempts
    self.extras = list(extra_storage_types) if extra_storage_types else []
    self.obs_dtype = observation_dtype
    self.term_dtype = terminal_dtype
    self.action_shape = action_shape
    self.action_dtype = action_dtype
    self.reward_shape = reward_shape
    self.reward_dtype = reward_dtype
    self.np_rng = np_rng if np_rng is not None else np.random
    self.store = {
        e.name: np.empty([replay_capacity] + list(e.shape), dtype=e.type)
        for e in self.storage_signature()
    }
    self.add_count = np.array(0)
    self.invalid_range = np.zeros((stack_size))
    self.discounts = np.array(
        [math.pow(gamma, k) for k in range(update_horizon)], dtype=np.float32)

  # -- signatures -------------------------------------------------------------
  def storage_signature(self):
    sig = [
        Element('observation', self.obs_shape, self.obs_dtype),
        Element('action', self.action_shape, self.action_dtype),
        Element('reward', self.reward_shape, self.reward_dtype),
        Element('terminal', (), self.term_dtype),
    ]
    return sig + list(self.extras)

  def add_signature(self):
    return self.storage_signature()

  def transition_signature(self, batch_size=None):
    b = self.batch if batch_size is None else batch_size
    state_shape = self.obs_shape + (self.stack,)
    sig = [
        Element('state', (b,) + state_shape, self.obs_dtype),
        Element('action', (b,) + self.action_shape, self.action_dtype),
        Element('reward', (b,) + self.reward_shape, self.reward_dtype),
        Element('next_state', (b,) + state_shape, self.obs_dtype),
        Element('next_action', (b,) + self.action_shape, self.action_dtype),
        Element('next_reward', (b,) + self.reward_shape, self.reward_dtype),
        Element('terminal', (b,), self.term_dtype),
        Element('indices', (b,), np.int32),
    ]
    for e in self.extras:
      sig.append(Element(e.name, (b,) + tuple(e.shape), e.type))
    return sig

  # Reference-API spellings, so shared tests can drive the port and the CUDA
  # classes through the same calls.
  def get_storage_signature(self):
    return self.storage_signature()

  def get_add_args_signature(self):
    return self.add_signature()

  def get_transition_elements(self, batch_size=None):
    return self.transition_signature(batch_size)

  def get_observation_stack(self, index):
    return self.observation_stack(index)

  def get_terminal_stack(self, index):
    return self.terminal_stack(index)

  def get_range(self, array, start_index, end_index):
    return self.span(array, start_index, end_index)

  def _check_add_types(self, *args):
    return self._check_shapes(*args)

  @property
  def _observation_shape(self):
    return self.obs_shape

  @property
  def _terminal_dtype(self):
    return self.term_dtype

  # -- bookkeeping --------------------------------------------------------------
  def cursor(self):
    return self.add_count % self.capacity

  def is_full(self):
    return self.add_count >= self.capacity

  def is_empty(self):
    return self.add_count == 0

  # -- add path (circular_replay_buffer.py:225-324) ------------------------------
  def _check_shapes(self, *args):
    sig = self.add_signature()
    if len(args) != len(sig):
      raise ValueError('Add expects {} elements, received {}'.format(
          len(sig), len(args)))
    for value, element in zip(args, sig):
      if isinstance(value, np.ndarray):
        shape = value.shape
      elif isinstance(value, (tuple, list)):
        shape = np.array(value).shape
      else:
        shape = tuple()
      if shape != tuple(element.shape):
        raise ValueError('arg has shape {}, expected {}'.format(
            shape, tuple(element.shape)))

  def add(self, observation, action, reward, terminal, *args):
    self._check_shapes(observation, action, reward, terminal, *args)
    if self.is_empty() or self.store['terminal'][self.cursor() - 1] == 1:
      for _ in range(self.stack - 1):
        self._write(*[np.zeros(e.shape, dtype=e.type)
                      for e in self.add_signature()])
    self._write(observation, action, reward, terminal, *args)

  def _write(self, *args):
    sig = self.add_signature()
    if len(args) != len(sig):
      raise ValueError('Add expects {} elements, received {}'.format(
          len(sig), len(args)))
    self._commit({e.name: v for e, v in zip(sig, args)})

  def _commit(self, row):
    slot = self.cursor()
    for name, value in row.items():
      self.store[name][slot] = value
    self.add_count += 1
    self.invalid_range = cursor_window(self.cursor(), self.capacity, self.stack,
                                       self.horizon)

  # -- reads (circular_replay_buffer.py:338-414) ---------------------------------
  def span(self, array, start, end):
    assert end > start, 'end_index must be larger than start_index'
    assert end >= 0
    assert start < self.capacity
    if not self.is_full():
      assert end <= self.cursor(), 'Index {} has not been added.'.format(start)
    if start % self.capacity < end % self.capacity:
      return array[start:end, ...]
    rows = [(start + k) % self.capacity for k in range(end - start)]
    return array[rows, ...]

  def observation_stack(self, index):
    frames = self.span(self.store['observation'], index - self.stack + 1,
                       index + 1)
    return np.moveaxis(frames, 0, -1)

  def terminal_stack(self, index):
    return self.span(self.store['terminal'], index - self.stack + 1, index + 1)

  def is_valid_transition(self, index):
    if index < 0 or index >= self.capacity:
      return False
    if not self.is_full():
      if index >= self.cursor() - self.horizon:
        return False
      if index < self.stack - 1:
        return False
    if index in set(self.invalid_range):
      return False
    if self.terminal_stack(index)[:-1].any():
      return False
    return True

  # -- sampling (circular_replay_buffer.py:436-558) -------------------------------
  def uniform_bounds(self):
    if self.is_full():
      lo = self.cursor() - self.capacity + self.stack - 1
      hi = self.cursor() - self.horizon
    else:
      lo = self.stack - 1
      hi = self.cursor() - self.horizon
      if hi <= lo:
        raise RuntimeError('Cannot sample a batch with fewer than stack size '
                           '({}) + update_horizon ({}) transitions.'.format(
                               self.stack, self.horizon))
    return lo, hi

  def sample_index_batch(self, batch_size):
    lo, hi = self.uniform_bounds()
    picked, misses = [], 0
    while len(picked) < batch_size and misses < self.max_attempts:
      candidate = self.np_rng.randint(lo, hi) % self.capacity
      if self.is_valid_transition(candidate):
        picked.append(candidate)
      else:
        misses += 1
    if len(picked) != batch_size:
      raise RuntimeError(
          'Max sample attempts: Tried {} times but only sampled {}'
          ' valid indices. Batch size is {}'.format(self.max_attempts,
                                                    len(picked), batch_size))
    return picked

  def sample_transition_batch(self, batch_size=None, indices=None):
    if batch_size is None:
      batch_size = self.batch
    if indices is None:
      indices = self.sample_index_batch(batch_size)
    assert len(indices) == batch_size
    sig = self.transition_signature(batch_size)
    out = tuple(np.empty(e.shape, dtype=e.type) for e in sig)
    for b, i in enumerate(indices):
      steps = [(i + k) % self.capacity for k in range(self.horizon)]
      flags = self.store['terminal'][steps]
      ends = flags.any()
      if ends:
        length = np.argmax(flags.astype(bool), 0) + 1
      else:
        length = self.horizon
      nxt = i + length
      rewards = self.span(self.store['reward'], i, nxt)
      for dst, e in zip(out, sig):
        if e.name == 'state':
          dst[b] = self.observation_stack(i)
        elif e.name == 'reward':
          dst[b] = np.sum(self.discounts[:length] * rewards, axis=0)
        elif e.name == 'next_state':
          dst[b] = self.observation_stack(nxt % self.capacity)
        elif e.name == 'next_action':
          dst[b] = self.store['action'][nxt % self.capacity]
        elif e.name == 'next_reward':
          dst[b] = self.store['reward'][nxt % self.capacity]
        elif e.name == 'terminal':
          dst[b] = ends
        elif e.name == 'indices':
          dst[b] = i
        elif e.name in self.store:
          dst[b] = self.store[e.name][i]
    return out


class PortPrioritizedReplay(PortReplay):
  """Prioritized replay (port of OutOfGraphPrioritizedReplayBuffer)."""

  def __init__(self, *args, **kwargs):
    py_rng = kwargs.pop('py_rng', None)
    super().__init__(*args, **kwargs)
    self.sum_tree = PortSumTree(self.capacity, rng=py_rng)

  def add_signature(self):
    return super().add_signature() + [Element('priority', (), np.float32)]

  def transition_signature(self, batch_size=None):
    b = self.batch if batch_size is None else batch_size
    return super().transition_signature(batch_size) + [
        Element('sampling_probabilities', (b,), np.float32)
    ]

  def _write(self, *args):
    sig = self.add_signature()
    if len(args) != len(sig):
      raise ValueError('Add expects {} elements, received {}'.format(
          len(sig), len(args)))
    row = {}
    priority = None
    for e, v in zip(sig, args):
      if e.name == 'priority':
        priority = v
      else:
        row[e.name] = v
    # prioritized_replay_buffer.py:139-140 — tree first, then the row.
    self.sum_tree.set(self.cursor(), priority)
    self._commit(row)

  def sample_index_batch(self, batch_size):
    picked = self.sum_tree.stratified_sample(batch_size)
    budget = self.max_attempts
    for slot in range(len(picked)):
      if self.is_valid_transition(picked[slot]):
        continue
      if budget == 0:
        raise RuntimeError(
            'Max sample attempts: Tried {} times but only sampled {}'
            ' valid indices. Batch size is {}'.format(self.max_attempts, slot,
                                                      batch_size))
      candidate = picked[slot]
      while not self.is_valid_transition(candidate) and budget > 0:
        candidate = self.sum_tree.sample()
        budget -= 1
      picked[slot] = candidate
    return picked

  def sample_transition_batch(self, batch_size=None, indices=None):
    out = super().sample_transition_batch(batch_size, indices)
    names = [e.name for e in self.transition_signature(batch_size)]
    where = out[names.index('indices')]
    out[names.index('sampling_probabilities')][:] = self.get_priority(where)
    return out

  def set_priority(self, indices, priorities):
    assert indices.dtype == np.int32, (
        'Indices must be integers, given: {}'.format(indices.dtype))
    for i, p in zip(indices, priorities):
      self.sum_tree.set(i, p)

  def get_priority(self, indices):
    assert indices.shape, 'Indices must be an array.'
    assert indices.dtype == np.int32, (
        'Indices must be int32s, given: {}'.format(indices.dtype))
    out = np.empty((len(indices)), dtype=np.float32)
    for k, i in enumerate(indices):
      out[k] = self.sum_tree.get(i)
    return out
