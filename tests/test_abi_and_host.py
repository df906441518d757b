"""CPU-side checks: the C-ABI library builds/loads and exports every symbol the
header declares, fails loudly without a GPU, and the host-side logic that needs no
device behaves like the reference."""
import ctypes
import os
import re

import numpy as np
import pytest

from dopamine_b200 import _native

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
  text = open(os.path.join(ROOT, 'include', 'b200_replay.h')).read()
  text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
  return sorted(set(re.findall(r'\b(b2r_[a-z0-9_]+)\s*\(', text)))


def test_library_exports_every_declared_symbol():
  lib = _native.lib()
  names = _declared_symbols()
  assert len(names) >= 40
  for name in names:
    assert hasattr(lib, name), name
  assert sorted(_native.SIGNATURES) == names
  assert lib.b2r_abi_version() == _native.ABI_VERSION == 3


def test_struct_layouts_match_the_header(tmp_path):
  """ctypes mirrors vs what a C compiler makes of include/b200_replay.h."""
  import subprocess
  src = tmp_path / 'sizes.c'
  src.write_text(
      '#include <stdio.h>\n#include <stddef.h>\n#include "b200_replay.h"\n'
      'int main(void) { printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\\n", '
      'sizeof(b2r_config), '
      'sizeof(b2r_batch), sizeof(b2r_c51_args), sizeof(b2r_trainer_config), '
      'offsetof(b2r_c51_args, min_probability), '
      'offsetof(b2r_trainer_config, seed), sizeof(b2r_iqn_args), '
      'offsetof(b2r_iqn_args, action_quantile_values), '
      'offsetof(b2r_iqn_args, next_action), sizeof(b2r_dqn_args)); return 0; }\n')
  exe = tmp_path / 'sizes'
  subprocess.check_call(['gcc', '-I', os.path.join(ROOT, 'include'), str(src),
                         '-o', str(exe)])
  got = [int(x) for x in subprocess.check_output([str(exe)]).split()]
  assert got == [ctypes.sizeof(_native.Config), ctypes.sizeof(_native.Batch),
                 ctypes.sizeof(_native.C51Args),
                 ctypes.sizeof(_native.TrainerConfig),
                 _native.C51Args.min_probability.offset,
                 _native.TrainerConfig.seed.offset, ctypes.sizeof(_native.IqnArgs),
                 _native.IqnArgs.action_quantile_values.offset,
                 _native.IqnArgs.next_action.offset, ctypes.sizeof(_native.DqnArgs)]
  assert ctypes.sizeof(_native.Config) == 96
  assert ctypes.sizeof(_native.Batch) == 8 * 8 + 8 * _native.MAX_EXTRAS + 8
  assert ctypes.sizeof(_native.C51Args) == 16 + 15 * 8


def test_no_cpu_fallback_without_a_device():
  import torch
  if torch.cuda.is_available():
    pytest.skip('a CUDA device is present')
  handle = ctypes.c_void_p()
  status = _native.lib().b2r_tree_create(16, ctypes.byref(handle))
  assert status == _native.ERR_CUDA
  assert 'no CPU fallback' in _native.last_error()
  from dopamine_b200.replay_memory import sum_tree
  with pytest.raises(_native.NativeError, match='no CPU fallback'):
    sum_tree.SumTree(16)


def test_argument_validation_needs_no_device():
  from dopamine_b200.replay_memory import circular_replay_buffer as crb
  from dopamine_b200.replay_memory import sum_tree
  with pytest.raises(ValueError, match='Sum tree capacity should be positive'):
    sum_tree.SumTree(-1)
  with pytest.raises(AssertionError):
    crb.OutOfGraphReplayBuffer(84, 4, 5, 32)
  with pytest.raises(ValueError, match='There is not enough capacity'):
    crb.OutOfGraphReplayBuffer((84, 84), 10, 10, 32)
  assert list(crb.invalid_range(6, 10, 4, 1)) == [5, 6, 7, 8, 9]
  assert list(crb.invalid_range(9, 10, 4, 1)) == [8, 9, 0, 1, 2]
  assert list(crb.invalid_range(0, 10, 4, 1)) == [9, 0, 1, 2, 3]
  assert list(crb.invalid_range(6, 10, 4, 3)) == [3, 4, 5, 6, 7, 8, 9]


def test_product_never_imports_the_oracle():
  pkg = os.path.join(ROOT, 'dopamine_b200')
  for base, _, files in os.walk(pkg):
    for f in files:
      if f.endswith(('.py', '.cu', '.cuh')):
        text = open(os.path.join(base, f)).read()
        assert 'import oracle' not in text and 'from oracle' not in text, f
        assert 'fast_oracle' not in text, f


def test_fastcall_shim_builds_and_binds():
  """csrc/fastcall.c: the CPython shim loads, binds the library's entry points and
  rejects an observation of the wrong size without touching the library."""
  fast = _native.fast()
  assert callable(fast.add_atari) and callable(fast.trainer_step)
  obs = np.zeros((84, 84), dtype=np.uint8)
  # wrong byte count -> -1 (caller falls back to the general path); handle unused
  assert fast.add_atari(0, 10, obs, 1, 0.5, 0, 1.0, 0, -1) == -1
  # not a buffer at all -> -1 as well
  assert fast.add_atari(0, 7056, object(), 1, 0.5, 0, 1.0, 0, -1) == -1
  with pytest.raises(TypeError):
    fast.add_atari(0, 7056, obs)


def test_bench_steps_per_graph_times_exactly_the_requested_steps():
  """bench.py captures several steps per CUDA graph launch; the group size must
  divide --steps so that exactly K steps are timed, whatever K the driver passes."""
  import bench
  for steps in list(range(1, 60)) + [97, 100, 2000, 20000]:
    for limit in (1, 10, 20):
      g = bench.steps_per_graph(steps, limit)
      assert 1 <= g <= limit and steps % g == 0
      assert all(steps % d for d in range(g + 1, min(limit, steps) + 1))
  assert bench.steps_per_graph(20000, 10) == 10 and bench.steps_per_graph(7, 10) == 7
