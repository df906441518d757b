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
