"""Import shim for the UNMODIFIED reference replay code (authoring container only).

TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only `oracle/make_golden.py`, the
`-m "not gpu"` tests and the CPU legs of `bench.py` (`--impl reference`,
`cpu_baseline`) may use this.  `/root/reference` does not exist on the GPU box; there
bench.py finds the copy of the three replay files that `make -C oracle _ref` made
(oracle/Makefile), the tests skip.  It lets the reference's
`dopamine/replay_memory/{sum_tree,circular_replay_buffer,prioritized_replay_buffer}.py`
be imported without TensorFlow / gin installed, by registering stub modules:

* `tensorflow`: only `tf.logging.info` is touched by the OutOfGraph* classes at
  construction time (circular_replay_buffer.py:146-156); `tf.gfile` only by
  save/load, which the shim maps onto the local filesystem so checkpoints written
  by the reference can be produced as fixtures.
* `gin` / `gin.tf`: `@gin.configurable(...)` is used as a decorator on the
  Wrapped* classes (circular_replay_buffer.py:690, prioritized_replay_buffer.py:255).

`load_reference_agents()` additionally imports `dopamine/agents/dqn/dqn_agent.py` and
`dopamine/agents/rainbow/rainbow_agent.py`.  Their module bodies touch `tf.contrib.slim`,
`tf.train.RMSPropOptimizer(...)` (a default argument) and import gym / atari_py / cv2
through `atari_lib`; a permissive stand-in object answers those.  The agents' GRAPH
code cannot run this way — only their plain-Python methods can (`begin_episode`, `step`,
`end_episode`, `_select_action`, `_train_step`, `_record_observation`, `_reset_state`,
`_store_transition`, `linearly_decaying_epsilon`), which is what
`oracle/make_golden.py:golden_actor` and `tests/test_oracle_golden.py` execute, bound to
a hand-made object in place of a constructed agent.

No reference source is copied or modified.
"""
import os
import sys
import types

_TRAVELLING_COPY = os.path.join(os.path.dirname(os.path.abspath(__file__)), '_ref')



def _stub_module(name):
  """An empty module with a real __spec__: importlib.util.find_spec (torch._dynamo probes
  'tensorflow' with it) raises on modules in sys.modules whose __spec__ is None."""
  import importlib.machinery
  mod = types.ModuleType(name)
  mod.__spec__ = importlib.machinery.ModuleSpec(name, None)
  return mod


def _default_root():
  """/root/reference in the authoring container; on the GPU box the byte-for-byte
  copy of the three replay files that `make -C oracle _ref` made (git-ignored, shipped
  by gpurun) — enough for load_reference(), not for load_reference_agents()."""
  if os.path.isdir('/root/reference/dopamine/replay_memory'):
    return '/root/reference'
  if os.path.isdir(os.path.join(_TRAVELLING_COPY, 'dopamine', 'replay_memory')):
    return _TRAVELLING_COPY
  return '/root/reference'


REFERENCE_ROOT = os.environ.get('DOPAMINE_REFERENCE_ROOT') or _default_root()


def reference_available():
  return os.path.isdir(os.path.join(REFERENCE_ROOT, 'dopamine', 'replay_memory'))


def _install_stubs():
  if 'tensorflow' not in sys.modules:
    tf = _stub_module('tensorflow')
    tf.logging = types.SimpleNamespace(
        info=lambda *a, **k: None, warning=lambda *a, **k: None)

    class _NotFoundError(Exception):

      def __init__(self, node_def=None, op=None, message=''):
        super().__init__(message)

    tf.errors = types.SimpleNamespace(NotFoundError=_NotFoundError)

    def _remove(path):
      try:
        os.remove(path)
      except FileNotFoundError:
        raise _NotFoundError(None, None, path)

    tf.gfile = types.SimpleNamespace(
        Exists=os.path.exists, Open=open, Remove=_remove)
    sys.modules['tensorflow'] = tf
  if 'gin' not in sys.modules:
    gin = _stub_module('gin')

    def configurable(*args, **kwargs):
      if len(args) == 1 and callable(args[0]) and not kwargs:
        return args[0]
      return lambda obj: obj

    gin.configurable = configurable
    gin_tf = _stub_module('gin.tf')
    gin.tf = gin_tf
    sys.modules['gin'] = gin
    sys.modules['gin.tf'] = gin_tf


def load_reference():
  """Returns (sum_tree, circular_replay_buffer, prioritized_replay_buffer)."""
  if not reference_available():
    raise RuntimeError('reference tree not present at ' + REFERENCE_ROOT)
  _install_stubs()
  if REFERENCE_ROOT not in sys.path:
    sys.path.insert(0, REFERENCE_ROOT)
  from dopamine.replay_memory import sum_tree  # pylint: disable=g-import-not-at-top
  from dopamine.replay_memory import circular_replay_buffer
  from dopamine.replay_memory import prioritized_replay_buffer
  return sum_tree, circular_replay_buffer, prioritized_replay_buffer


class _Anything(object):
  """Answers any attribute access or call with itself (import-time stand-in)."""

  def __getattr__(self, name):
    if name.startswith('__') and name.endswith('__'):
      raise AttributeError(name)
    return self

  def __call__(self, *args, **kwargs):
    return self


def _module_fallback(anything):
  def fallback(name):
    if name.startswith('__') and name.endswith('__'):
      raise AttributeError(name)
    return anything
  return fallback


def load_reference_agents():
  """Returns the reference's (dqn_agent, rainbow_agent) modules; see the header."""
  load_reference()
  anything = _Anything()
  tf = sys.modules['tensorflow']
  if not isinstance(tf, types.ModuleType) or getattr(tf, '__file__', None):
    raise RuntimeError('a real TensorFlow is installed: import the agents directly')
  if '__getattr__' not in tf.__dict__:
    tf.__getattr__ = _module_fallback(anything)  # PEP 562: tf.contrib, tf.train, ...
  for name in ('atari_py', 'gym', 'gym.spaces', 'gym.spaces.box', 'cv2'):
    if name not in sys.modules:
      mod = _stub_module(name)
      mod.__getattr__ = _module_fallback(anything)
      sys.modules[name] = mod
  gin = sys.modules['gin']
  if '__getattr__' not in gin.__dict__:
    gin.__getattr__ = _module_fallback(anything)  # gin.constant, gin.REQUIRED, ...
  from dopamine.agents.dqn import dqn_agent  # pylint: disable=g-import-not-at-top
  from dopamine.agents.rainbow import rainbow_agent
  return dqn_agent, rainbow_agent
