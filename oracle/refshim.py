"""Import shim for the UNMODIFIED reference replay code (authoring container only).

TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only `oracle/make_golden.py` and the
`-m "not gpu"` tests may use this, and only when `/root/reference` exists (it
does not exist on the GPU box).  It lets the reference's
`dopamine/replay_memory/{sum_tree,circular_replay_buffer,prioritized_replay_buffer}.py`
be imported without TensorFlow / gin installed, by registering stub modules:

* `tensorflow`: only `tf.logging.info` is touched by the OutOfGraph* classes at
  construction time (circular_replay_buffer.py:146-156); `tf.gfile` only by
  save/load, which the shim maps onto the local filesystem so checkpoints written
  by the reference can be produced as fixtures.
* `gin` / `gin.tf`: `@gin.configurable(...)` is used as a decorator on the
  Wrapped* classes (circular_replay_buffer.py:690, prioritized_replay_buffer.py:255).

No reference source is copied or modified.
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get('DOPAMINE_REFERENCE_ROOT', '/root/reference')


def reference_available():
  return os.path.isdir(os.path.join(REFERENCE_ROOT, 'dopamine', 'replay_memory'))


def _install_stubs():
  if 'tensorflow' not in sys.modules:
    tf = types.ModuleType('tensorflow')
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
    gin = types.ModuleType('gin')

    def configurable(*args, **kwargs):
      if len(args) == 1 and callable(args[0]) and not kwargs:
        return args[0]
      return lambda obj: obj

    gin.configurable = configurable
    gin_tf = types.ModuleType('gin.tf')
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
