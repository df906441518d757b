"""The Rainbow learner on top of the replay path (BASELINE config 5, SURVEY 8f rank
2): network shapes, one full train step against the oracles, target sync."""
import numpy as np
import pytest

from oracle import c51_port


def test_same_padding_matches_tensorflow_shapes():
  from dopamine_b200.agents.rainbow import agent
  # atari_lib.py:126-131 with SAME padding: 84 -> 21 -> 11 -> 11
  assert agent._same_pad(84, 8, 4) == (2, 2)
  assert agent._same_pad(21, 4, 2) == (1, 2)
  assert agent._same_pad(11, 3, 1) == (1, 1)
  assert agent._same_pad(5, 1, 1) == (0, 0)


def test_optimizer_is_tensorflow_adam_not_torch_adam():
  """rainbow_agent.py:69-71 + rainbow.gin:21-25: tf.train.AdamOptimizer(6.25e-5,
  epsilon=1.5e-4).  TensorFlow 1.x: lr_t = lr sqrt(1 - b2^t) / (1 - b1^t), theta -= lr_t
  m / (sqrt(v) + eps), eps on the UNcorrected sqrt(v) — restated here in numpy float32,
  op by op.  torch.optim.Adam with the same numbers takes a 20x larger first step."""
  import torch
  from dopamine_b200.agents.rainbow import agent
  rng = np.random.RandomState(0)
  shapes = [(7, 5), (11,), (3, 2, 4)]
  w0 = [rng.randn(*sh).astype(np.float32) for sh in shapes]
  params = [torch.nn.Parameter(torch.tensor(w.copy())) for w in w0]
  opt = agent.make_tf_adam(params, lr=6.25e-5, epsilon=1.5e-4)
  ref = torch.optim.Adam([torch.nn.Parameter(torch.tensor(w.copy())) for w in w0],
                         lr=6.25e-5, eps=1.5e-4)
  w = [x.copy() for x in w0]
  m = [np.zeros_like(x) for x in w0]
  v = [np.zeros_like(x) for x in w0]
  b1, b2, eps = np.float32(0.9), np.float32(0.999), np.float32(1.5e-4)
  for t in range(1, 8):
    grads = [(1e-3 * rng.randn(*sh)).astype(np.float32) for sh in shapes]
    for p, g in zip(params, grads):
      p.grad = torch.tensor(g.copy())
    for p, g in zip(ref.param_groups[0]['params'], grads):
      p.grad = torch.tensor(g.copy())
    opt.step()
    ref.step()
    lr_t = np.float32(6.25e-5 * np.sqrt(1.0 - 0.999 ** t) / (1.0 - 0.9 ** t))
    for k, g in enumerate(grads):
      m[k] = b1 * m[k] + (np.float32(1) - b1) * g
      v[k] = b2 * v[k] + (np.float32(1) - b2) * g * g
      w[k] = w[k] - lr_t * m[k] / (np.sqrt(v[k]) + eps)
    for p, x, x0 in zip(params, w, w0):
      moved = np.abs(x - x0).max()
      np.testing.assert_allclose(p.detach().numpy() - x0, x - x0, rtol=2e-5,
                                 atol=1e-6 * moved)
    if t == 1:  # the two optimizers are NOT interchangeable at Rainbow's epsilon
      ours = np.abs(params[0].detach().numpy() - w0[0]).max()
      torch_adam = np.abs(ref.param_groups[0]['params'][0].detach().numpy() - w0[0]).max()
      assert torch_adam > 3.0 * ours


@pytest.fixture(scope='module')
def learner_mod():
  import torch
  if not torch.cuda.is_available():
    pytest.fail('-m gpu tests need a CUDA device (no CPU fallback exists)')
  from dopamine_b200.agents.rainbow import agent
  return agent


@pytest.mark.gpu
def test_network_shapes_and_init(learner_mod):
  import torch
  net = learner_mod.make_rainbow_network(18, 51).cuda()
  x = torch.randint(0, 256, (5, 84, 84, 4), dtype=torch.uint8, device='cuda')
  out = net(x)
  assert tuple(out.shape) == (5, 18, 51)
  assert net.fc1.in_features == 11 * 11 * 64  # 7 744 features
  assert sum(p.numel() for p in net.parameters()) == (
      32 * 4 * 64 + 32 + 64 * 32 * 16 + 64 + 64 * 64 * 9 + 64 +
      7744 * 512 + 512 + 512 * 918 + 918)
  limit = (3.0 / np.sqrt(3.0) / (4 * 64)) ** 0.5  # first conv, fan-in 4*8*8
  w = net.convs[0].weight
  assert float(w.abs().max()) <= limit and float(w.abs().max()) > 0.9 * limit


def _numpy_rainbow_network(net, state, num_actions, num_atoms):
  """atari_lib.rainbow_network (atari_lib.py:108-144) restated in numpy float64 over the
  torch module's weights: cast + / 255 (124-125), three SAME-padded convolutions with
  ReLU (126-131; TensorFlow pads floor(p / 2) before and the rest after), flatten in
  NHWC order (132), fully connected 512 + ReLU (133-135), fully connected A * N
  (136-140), reshape to (B, A, N) (141)."""
  x = state.astype(np.float64) / 255.0  # (B, H, W, C)
  specs = [(8, 4), (4, 2), (3, 1)]
  for conv, (k, stride) in zip(net.convs, specs):
    w = conv.weight.detach().cpu().numpy().astype(np.float64)  # (O, I, kh, kw)
    b = conv.bias.detach().cpu().numpy().astype(np.float64)
    size = x.shape[1]
    out = -(-size // stride)
    total = max((out - 1) * stride + k - size, 0)
    lo, hi = total // 2, total - total // 2
    x = np.pad(x, ((0, 0), (lo, hi), (lo, hi), (0, 0)))
    windows = np.lib.stride_tricks.sliding_window_view(x, (k, k), axis=(1, 2))
    windows = windows[:, ::stride, ::stride]  # (B, out, out, C, kh, kw)
    x = np.maximum(np.einsum('bhwikl,oikl->bhwo', windows, w) + b, 0.0)
  x = x.reshape(x.shape[0], -1)  # NHWC flatten
  w1 = net.fc1.weight.detach().cpu().numpy().astype(np.float64)
  x = np.maximum(x @ w1.T + net.fc1.bias.detach().cpu().numpy(), 0.0)
  w2 = net.fc2.weight.detach().cpu().numpy().astype(np.float64)
  x = x @ w2.T + net.fc2.bias.detach().cpu().numpy()
  return x.reshape(-1, num_actions, num_atoms)


@pytest.mark.gpu
def test_network_forward_matches_the_reference_architecture(learner_mod):
  """The cuDNN network against a numpy restatement of atari_lib.py:108-144 (layer order,
  SAME padding sides, NHWC flatten order, ReLUs, output reshape) in float32 arithmetic
  (TF-1.x has no TF32: it is switched off for the comparison)."""
  import torch
  torch.manual_seed(4)
  net = learner_mod.make_rainbow_network(6, 51).cuda()
  with torch.no_grad():
    for p in net.parameters():  # biases are zero-initialised: give them values
      if p.dim() == 1:
        p.uniform_(-0.05, 0.05)
  rng = np.random.RandomState(2)
  state = rng.randint(0, 256, size=(3, 84, 84, 4)).astype(np.uint8)
  tf32 = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
  torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
  try:
    with torch.no_grad():
      got = net(torch.as_tensor(state, device='cuda')).cpu().numpy()
  finally:
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf32
  want = _numpy_rainbow_network(net, state, 6, 51)
  np.testing.assert_allclose(got, want, rtol=1e-4, atol=1e-5)


@pytest.mark.gpu
@pytest.mark.parametrize('shape', [(5, 84, 84, 4), (2, 7, 6, 4), (3, 5, 5, 3), (1, 84, 84, 1)])
def test_network_input_kernel_is_cast_and_division_bit_for_bit(learner_mod, shape):
  """b2r_stack_to_planes_device: tf.cast(state, tf.float32), tf.div(net, 255.)
  (atari_lib.py:124-125) and the NCHW layout in one pass — float32 equal to the eager
  tensor ops bit for bit, float16 equal to that quotient rounded once."""
  import torch
  rng = np.random.RandomState(1)
  host = rng.randint(0, 256, size=shape).astype(np.uint8)

  def expected(x):  # numpy: a true float32 division (torch's div by a scalar multiplies
    # by the reciprocal, which differs in the last bit for some bytes)
    return np.ascontiguousarray(
        np.transpose(x, (0, 3, 1, 2)).astype(np.float32) / np.float32(255.))

  state = torch.as_tensor(host, device='cuda')
  want = expected(host)
  got = learner_mod.network_input(state)
  assert tuple(got.shape) == want.shape and got.dtype == torch.float32
  assert got.cpu().numpy().tobytes() == want.tobytes()
  half = learner_mod.network_input(state, half=True)
  assert half.dtype == torch.float16
  assert half.cpu().numpy().tobytes() == want.astype(np.float16).tobytes()
  # every byte value
  ramp = np.tile(np.arange(256, dtype=np.uint8), 4).reshape(1, 16, 16, 4)
  got = learner_mod.network_input(torch.as_tensor(ramp, device='cuda'))
  assert got.cpu().numpy().tobytes() == expected(ramp).tobytes()


@pytest.mark.gpu
def test_train_step_updates_network_and_priorities(learner_mod):
  """One update: the priorities written back are sqrt(loss + 1e-10) of the C51 loss
  of THIS batch under the networks' logits (numpy restatement), the online network
  moves, the target network does not until sync_target()."""
  import torch
  rng = np.random.RandomState(0)
  learner = learner_mod.RainbowLearner(6, replay_capacity=2000, batch_size=16, seed=3)
  for k in range(600):
    learner.store_transition(rng.randint(0, 256, size=(84, 84)).astype(np.uint8),
                             int(rng.randint(6)), float(np.clip(rng.randn(), -1, 1)),
                             int(rng.rand() < 0.02))
  mem = learner.memory
  before = [p.detach().clone() for p in learner.online.parameters()]
  target_before = [p.detach().clone() for p in learner.target.parameters()]
  loss = learner.train_step()
  torch.cuda.synchronize()
  assert np.isfinite(float(loss))
  # the batch that was just used is still in the reused output buffers
  _, arrays, _ = mem._alloc_outputs(16, True)
  state, action, reward, next_state, _, _, terminal, indices, probs = arrays[:9]
  with torch.no_grad():
    # the networks as they were BEFORE the optimizer step
    old = learner_mod.make_rainbow_network(6, 51).cuda()
    old.load_state_dict({k: v for (k, _), v in zip(
        learner.online.state_dict().items(), before)})
    online_logits = old(state).cpu().numpy()
    target_logits = learner.target(next_state).cpu().numpy()
  ref = c51_port.rainbow_update(reward.cpu().numpy(), terminal.cpu().numpy(),
                                action.cpu().numpy(), probs.cpu().numpy(),
                                online_logits, target_logits, update_horizon=3)
  want_loss = float(np.mean(ref['weights'] * ref['loss']))
  assert abs(float(loss) - want_loss) <= 2e-5 * max(1.0, abs(want_loss))
  got_prio = mem.get_priority(indices).cpu().numpy()
  idx = indices.cpu().numpy()
  last = {int(i): k for k, i in enumerate(idx)}  # later duplicates win
  for i, k in last.items():
    np.testing.assert_allclose(got_prio[k], ref['priorities'][k], rtol=2e-5, atol=1e-6)
  assert any(not torch.equal(a, b) for a, b in
             zip(before, learner.online.parameters()))
  assert all(torch.equal(a, b) for a, b in
             zip(target_before, learner.target.parameters()))
  learner.sync_target()
  assert all(torch.equal(a, b) for a, b in
             zip(learner.online.parameters(), learner.target.parameters()))
  # cadence: update every 4th call (dqn_agent.py:418-442)
  done = sum(learner.step_cadence() is not None for _ in range(12))
  assert done == 3


@pytest.mark.gpu
def test_learner_reduces_loss_on_a_fixed_replay(learner_mod):
  """Sanity of the whole loop: repeated updates on a frozen replay lower the loss."""
  import torch
  rng = np.random.RandomState(1)
  learner = learner_mod.RainbowLearner(4, replay_capacity=1000, batch_size=32,
                                       seed=5, learning_rate=1e-3,
                                       replay_scheme='uniform')
  for k in range(400):
    learner.store_transition(rng.randint(0, 256, size=(84, 84)).astype(np.uint8),
                             int(rng.randint(4)), float(np.clip(rng.randn(), -1, 1)),
                             int(rng.rand() < 0.02))
  first = np.mean([float(learner.train_step()) for _ in range(5)])
  for _ in range(60):
    learner.train_step()
  last = np.mean([float(learner.train_step()) for _ in range(5)])
  torch.cuda.synchronize()
  assert last < first


@pytest.mark.gpu
def test_graph_replayed_learner_equals_eager(learner_mod):
  """cuda_graph=True replays the whole update as one CUDA graph; with adds between
  updates it must produce the same parameters, priorities and losses as the eager
  learner fed identically."""
  import torch
  rng = np.random.RandomState(2)
  frames = rng.randint(0, 256, size=(700, 84, 84)).astype(np.uint8)
  acts = rng.randint(0, 5, size=700)
  rews = np.clip(rng.randn(700), -1, 1)
  terms = rng.rand(700) < 0.02

  def run(cuda_graph):
    torch.manual_seed(0)
    learner = learner_mod.RainbowLearner(5, replay_capacity=1000, batch_size=16,
                                         seed=7, cuda_graph=cuda_graph)
    k = 0
    for _ in range(500):
      learner.store_transition(frames[k], int(acts[k]), float(rews[k]), int(terms[k]))
      k += 1
    losses = []
    for it in range(12):
      for _ in range(4):
        learner.store_transition(frames[k], int(acts[k]), float(rews[k]), int(terms[k]))
        k += 1
      if not cuda_graph:
        if it == 0:
          for _ in range(3):  # the graphed learner warms up with 3 updates at capture
            learner.train_step()
        else:
          # the captured sampler keeps the host draw offset of the capture call (4)
          # and advances only the device counter: give the eager run the same stream
          learner.memory._draw_counter = 3
      losses.append(float(learner.train_step().detach()))
    torch.cuda.synchronize()
    params = [p.detach().cpu().numpy().copy() for p in learner.online.parameters()]
    nodes = learner.memory.sum_tree.nodes
    return losses, params, nodes

  eager = run(False)
  graphed = run(True)
  np.testing.assert_allclose(graphed[0], eager[0], rtol=2e-3)
  # Adam divides by sqrt(v): near-zero gradient entries amplify cuDNN's run-to-run
  # rounding differences, hence the absolute tolerance of a few learning rates
  for a, b in zip(graphed[1], eager[1]):
    np.testing.assert_allclose(a, b, rtol=2e-3, atol=5e-4)
  # same sampled indices and (up to cuDNN's algorithm choice) priorities
  for a, b in zip(graphed[2], eager[2]):
    np.testing.assert_allclose(a, b, rtol=5e-3, atol=1e-5)
