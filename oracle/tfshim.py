"""numpy stand-ins for the TensorFlow 1.x ops that the reference's LOSS-BUILDING methods
call.  TEST INFRASTRUCTURE, authoring container only.

With these installed on the stub `tensorflow` module of `oracle/refshim.py`, the
UNMODIFIED reference methods

  rainbow_agent.RainbowAgent._build_target_distribution / _build_train_op  (RA:200-305)
  rainbow_agent.project_distribution                                        (RA:340-494)
  dqn_agent.DQNAgent._build_networks / _build_target_q_op / _build_train_op (DQ:237-322)
  implicit_quantile_agent.ImplicitQuantileAgent._build_networks /
      _build_target_quantile_values_op / _build_train_op                    (IQ:120-321)

execute eagerly on float32 numpy arrays (bound to a hand-made object whose "networks"
return seeded arrays; `oracle/make_golden.py:golden_losses`).  What this pins is every
decision the reference's code makes — tiling, reshaping, transposes, gather indices,
which tensor is subtracted from which, masks, reduction axes, the order of the
elementwise operations.  What it does NOT reproduce is TensorFlow's own kernels: each op
below is numpy's float32 arithmetic with the op's documented definition (softmax and
softmax cross-entropy with the max-shift TF uses, `tf.losses.huber_loss` as TF 1.x
defines it).  DESIGN.md section 2 states the parity status accordingly.
"""
import contextlib
import types

import numpy as np

F32 = np.float32


class _Shape(tuple):
  """A tuple with the two TensorShape methods project_distribution calls (RA:384-386)."""

  def assert_is_compatible_with(self, other):
    other = tuple(other)
    if len(self) != len(other) or any(a != b for a, b in zip(self, other)):
      raise ValueError('Shapes %s and %s are incompatible' % (tuple(self), other))

  def assert_has_rank(self, rank):
    if len(self) != rank:
      raise ValueError('Shape %s must have rank %d' % (tuple(self), rank))


class Tensor(np.ndarray):
  """ndarray whose .shape answers like a TensorShape; arithmetic is numpy's."""

  @property
  def shape(self):
    return _Shape(np.ndarray.shape.__get__(self))


def tensor(x, dtype=None):
  return np.asarray(x, dtype=dtype).view(Tensor)


def _plain(x):
  return np.asarray(x)


def _reduce(fn, x, axis=None, reduction_indices=None, name=None, keepdims=False):
  del name
  if axis is None:
    axis = reduction_indices
  x = _plain(x)
  kw = dict(axis=axis, keepdims=keepdims)
  if fn in (np.sum, np.mean) and x.dtype == F32:
    kw['dtype'] = F32
  return tensor(fn(x, **kw))


def _softmax(logits, axis=-1):
  x = _plain(logits).astype(F32)
  e = np.exp(x - x.max(axis=axis, keepdims=True))
  return tensor((e / e.sum(axis=axis, keepdims=True, dtype=F32)).astype(F32))


def _softmax_cross_entropy_with_logits(labels=None, logits=None, **unused):
  """-sum(labels * log_softmax(logits)) over the last axis, log_softmax computed as
  (logits - max) - log(sum(exp(logits - max))) (TF's xent kernel)."""
  x = _plain(logits).astype(F32)
  t = _plain(labels).astype(F32)
  shifted = (x - x.max(axis=-1, keepdims=True)).astype(F32)
  lse = np.log(np.exp(shifted).sum(axis=-1, keepdims=True, dtype=F32)).astype(F32)
  return tensor(-(t * (shifted - lse).astype(F32)).sum(axis=-1, dtype=F32))


def _huber_loss(labels, predictions, weights=1.0, delta=1.0, scope=None,
                loss_collection=None, reduction=None):
  """tf.losses.huber_loss, TF 1.x: error = predictions - labels; quadratic =
  min(|error|, delta); linear = |error| - quadratic; 0.5 quadratic^2 + delta linear."""
  del scope, loss_collection
  assert reduction == 'none' and weights == 1.0
  error = (_plain(predictions).astype(F32) - _plain(labels).astype(F32)).astype(F32)
  abs_error = np.abs(error)
  quadratic = np.minimum(abs_error, F32(delta))
  linear = (abs_error - quadratic).astype(F32)
  return tensor((F32(0.5) * (quadratic * quadratic).astype(F32) +
                 (F32(delta) * linear).astype(F32)).astype(F32))


def _linspace(start, stop, num, name=None):
  """TF 1.x LinSpaceOp (tensorflow/core/kernels/sequence_ops.cc): in the output type,
  step = (stop - start) / (num - 1); out[i] = start + step * i (up to
  1 ulp away from np.linspace(...).astype(float32))."""
  del name
  start, stop = F32(start), F32(stop)
  step = F32((stop - start) / F32(num - 1))
  return tensor((start + step * np.arange(num, dtype=F32)).astype(F32))


def _gather_nd(params, indices):
  idx = _plain(indices)
  return tensor(_plain(params)[tuple(idx[..., k] for k in range(idx.shape[-1]))])


def _one_hot(indices, depth, on_value=1., off_value=0., name=None):
  del name
  idx = _plain(indices)
  out = np.full(idx.shape + (depth,), off_value, dtype=F32)
  np.put_along_axis(out, idx[..., None].astype(np.int64), F32(on_value), axis=-1)
  return tensor(out)


def _assert(condition, data, **unused):
  if not bool(np.all(_plain(condition))):
    raise ValueError('assertion failed: %r' % (data,))
  return None


@contextlib.contextmanager
def _scope(*unused_args, **unused_kwargs):
  yield


def install(tf):
  """Sets the numpy stand-ins on the stub module `tf` (never on a real TensorFlow)."""
  if getattr(tf, '__file__', None):
    raise RuntimeError('refusing to patch a real TensorFlow')
  tf.float32, tf.int32, tf.int64 = np.float32, np.int32, np.int64
  tf.cast = lambda x, dtype, name=None: tensor(_plain(x).astype(dtype))
  tf.to_float = lambda x, name=None: tensor(_plain(x).astype(F32))
  tf.to_int64 = lambda x, name=None: tensor(_plain(x).astype(np.int64))
  tf.shape = lambda x, name=None: np.array(np.shape(_plain(x)), dtype=np.int64)
  tf.size = lambda x, name=None: np.int64(np.size(_plain(x)))
  tf.tile = lambda x, multiples, name=None: tensor(
      np.tile(_plain(x), [int(m) for m in multiples]))
  tf.reshape = lambda x, shape, name=None: tensor(
      np.reshape(_plain(x), [int(s) for s in shape]))
  tf.transpose = lambda x, perm=None, name=None: tensor(np.transpose(_plain(x), perm))
  tf.squeeze = lambda x, axis=None, name=None: tensor(np.squeeze(_plain(x), axis=axis))
  tf.concat = lambda values, axis, name=None: tensor(
      np.concatenate([_plain(v) for v in values], axis=axis))
  tf.range = lambda *a, **k: tensor(np.arange(*[int(x) for x in a]))
  tf.abs = lambda x, name=None: tensor(np.abs(_plain(x)))
  tf.sqrt = lambda x, name=None: tensor(np.sqrt(_plain(x)))
  tf.equal = lambda a, b, name=None: tensor(_plain(a) == _plain(b))
  tf.clip_by_value = lambda x, lo, hi, name=None: tensor(
      np.clip(_plain(x), lo, hi).astype(_plain(x).dtype))
  tf.argmax = lambda x, axis=None, name=None: tensor(
      np.argmax(_plain(x), axis=axis).astype(np.int64))
  tf.reduce_sum = lambda x, *a, **k: _reduce(np.sum, x, *a, **k)
  tf.reduce_mean = lambda x, *a, **k: _reduce(np.mean, x, *a, **k)
  tf.reduce_max = lambda x, *a, **k: _reduce(np.max, x, *a, **k)
  tf.reduce_all = lambda x, *a, **k: _reduce(np.all, x, *a, **k)
  tf.linspace = _linspace
  tf.gather_nd = _gather_nd
  tf.one_hot = _one_hot
  tf.stop_gradient = lambda x, name=None: x
  tf.no_op = lambda name=None: None
  tf.Assert = _assert
  tf.control_dependencies = _scope
  tf.variable_scope = _scope
  tf.make_template = lambda name, fn, **unused: types.SimpleNamespace(name=name, fn=fn)
  tf.nn = types.SimpleNamespace(
      softmax=_softmax,
      softmax_cross_entropy_with_logits=_softmax_cross_entropy_with_logits)
  tf.losses = types.SimpleNamespace(
      huber_loss=_huber_loss, Reduction=types.SimpleNamespace(NONE='none'))
  tf.softmax = _softmax  # (what tf.contrib.layers.softmax computes on the last axis)
  return tf
