"""GPU: the one-kernel rollout policy (paintrl_policy_act, csrc/paintrl_policy.cuh; the MLP of paint_ppo.py:179-183)
against a plain PyTorch FP32 evaluation of the same weights, and its sampling against a NumPy replica of the
kernel's counter-based random numbers.

Tolerances: layer 2 runs on the tensor cores with BF16 inputs and FP32 accumulation, the activations use
tanh.approx -- logits / values within 2e-2 absolute of the FP32 reference (|logits| ~ 0.1 ... 1 with the random-init
weights); the sampled action must be the arg-max of (the kernel's own logits + Gumbel noise recomputed on the host)
except where the two best candidates are closer than 1e-4."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

_M = (1 << 64) - 1


def _mix(x):
    x = (x + np.uint64(0x9E3779B97F4A7C15))
    x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return x ^ (x >> np.uint64(31))


def _uniform(seed, env, ctr, k):
    """pol_uniform of paintrl_policy.cuh on uint64 arrays (wrap-around arithmetic)."""
    with np.errstate(over='ignore'):
        inner = _mix((env.astype(np.uint64) << np.uint64(32)) | np.uint64(ctr))
        h = _mix(np.uint64(seed) ^ inner ^ (np.uint64(k) * np.uint64(0xD6E8FEB86659FD93)))
    return ((h >> np.uint64(40)).astype(np.float32) + np.float32(0.5)) * np.float32(1.0 / 16777216.0)


def _reference(policy, obs):
    torch.backends.cuda.matmul.allow_tf32 = False
    h = obs.to(torch.float32)
    for w, b in policy.layers:
        h = torch.tanh(h @ w + b)
    return h @ policy.head_w + policy.head_b


@pytest.mark.parametrize('obs_dim,n_out,batch', [(6, 4, 1000), (16, 4, 128), (2, 8, 77), (32, 15, 4096)])
def test_discrete_policy_kernel(cuda_device, obs_dim, n_out, batch):
    from paintrl_b200.rollout import MlpPolicy
    pol = MlpPolicy(obs_dim, n_out, device=cuda_device, seed=5, discrete=True)
    for w in pol.weights():                        # biases are zero at init: give them values
        if w.dim() == 1:
            w.copy_(torch.linspace(-0.3, 0.3, w.numel(), device=cuda_device))
    pol.head_w.mul_(10.0)                          # logits of order 1
    assert pol.enable_native(batch)
    gen = torch.Generator(device=cuda_device); gen.manual_seed(3)
    obs = torch.rand(batch, obs_dim, generator=gen, device=cuda_device, dtype=torch.float64)
    ref = _reference(pol, obs).cpu().numpy()
    actions = torch.full((batch,), -1, dtype=torch.int64, device=cuda_device)
    logp = torch.zeros(batch, dtype=torch.float32, device=cuda_device)
    value = torch.zeros(batch, dtype=torch.float32, device=cuda_device)
    logits = torch.zeros(batch, n_out + 1, dtype=torch.float32, device=cuda_device)
    counts = np.zeros(n_out, dtype=np.int64)
    for draw in range(3):
        pol.native.act_into(obs, actions, logp, value, logits=logits)
        lg = logits.cpu().numpy()
        assert np.abs(lg - ref).max() < 2e-2, np.abs(lg - ref).max()
        assert np.array_equal(value.cpu().numpy(), lg[:, n_out])
        env = np.arange(batch)
        noisy = np.stack([lg[:, o] - np.log(-np.log(_uniform(5, env, draw, o))) for o in range(n_out)], axis=1)
        a = actions.cpu().numpy()
        assert a.min() >= 0 and a.max() < n_out
        top2 = np.sort(noisy, axis=1)[:, -2:]
        clear = (top2[:, 1] - top2[:, 0]) > 1e-4
        assert clear.mean() > 0.99
        assert np.array_equal(a[clear], noisy.argmax(axis=1)[clear])
        lse = np.log(np.exp(lg[:, :n_out] - lg[:, :n_out].max(axis=1, keepdims=True)).sum(axis=1)) + lg[:, :n_out].max(axis=1)
        assert np.allclose(logp.cpu().numpy(), lg[env, a] - lse, atol=1e-4)
        counts += np.bincount(a, minlength=n_out)
    if batch >= 1000:      # every action is drawn; consecutive draws differ (the counters advance)
        assert counts.min() > 0
    # bootstrap call: value only, no draw, counters untouched
    v2 = torch.zeros_like(value)
    pol.native.act_into(obs, None, None, v2, sample=False)
    assert torch.equal(v2, value)
    pol.native.act_into(obs, actions, logp, value, logits=logits)
    noisy = np.stack([logits.cpu().numpy()[:, o] - np.log(-np.log(_uniform(5, np.arange(batch), 3, o))) for o in range(n_out)], axis=1)
    top2 = np.sort(noisy, axis=1)[:, -2:]
    clear = (top2[:, 1] - top2[:, 0]) > 1e-4
    assert np.array_equal(actions.cpu().numpy()[clear], noisy.argmax(axis=1)[clear])
    pol.native.close()


def test_continuous_policy_kernel(cuda_device):
    from paintrl_b200.rollout import MlpPolicy
    batch, obs_dim, n_out = 513, 16, 2
    pol = MlpPolicy(obs_dim, n_out, device=cuda_device, seed=9, discrete=False)
    pol.head_w.mul_(10.0)
    assert pol.enable_native(batch)
    gen = torch.Generator(device=cuda_device); gen.manual_seed(4)
    obs = torch.rand(batch, obs_dim, generator=gen, device=cuda_device, dtype=torch.float64)
    ref = _reference(pol, obs).cpu().numpy()
    actions = torch.zeros(batch, n_out, dtype=torch.float64, device=cuda_device)
    logp = torch.zeros(batch, dtype=torch.float32, device=cuda_device)
    value = torch.zeros(batch, dtype=torch.float32, device=cuda_device)
    logits = torch.zeros(batch, n_out + 1, dtype=torch.float32, device=cuda_device)
    pol.native.act_into(obs, actions, logp, value, logits=logits)
    lg = logits.cpu().numpy()
    assert np.abs(lg - ref).max() < 2e-2
    env = np.arange(batch)
    z = np.stack([np.sqrt(-2 * np.log(_uniform(9, env, 0, 2 * o))) * np.cos(2 * np.pi * _uniform(9, env, 0, 2 * o + 1)) for o in range(n_out)], axis=1)
    assert np.allclose(actions.cpu().numpy(), np.tanh(lg[:, :n_out]) + z, atol=2e-3)
    assert np.allclose(logp.cpu().numpy(), (-0.5 * z * z - 0.9189385332046727).sum(axis=1), atol=2e-3)
    assert abs(z.mean()) < 0.15 and 0.8 < z.std() < 1.2
    pol.native.close()


def test_rollout_worker_uses_the_policy_kernel(cuda_device):
    """A fragment collected with the policy kernel inside a captured CUDA graph: the actions in the fragment are the ones
    the environments were stepped with (replayed through a second engine), and a replay draws new noise."""
    from paintrl_b200.batched_env import BatchedPaintEnv
    from paintrl_b200.config import EnvConfig
    from paintrl_b200.rollout import MlpPolicy, RolloutWorker
    from test_gpu_oracle_batch import BASE
    n, T = 256, 20
    cfg = EnvConfig(dict(BASE), auto_reset=True, seed=3)
    env = BatchedPaintEnv(n, cfg, device=cuda_device)
    twin = BatchedPaintEnv(n, cfg, device=cuda_device)
    pol = MlpPolicy(env.obs_dim, 4, device=cuda_device, seed=1)
    worker = RolloutWorker(env, pol, fragment_length=T, use_cuda_graph=True)
    assert pol.native is not None
    start = (np.arange(n) % env.n_starts).astype(np.int32)
    worker.start(start)
    twin.reset(start)
    seen = []
    for it in range(3):                      # eager, capture, replay
        launches0 = env.stats()['kernel_launches']
        f, stats = worker.collect()
        acts = f.actions.clone()
        seen.append(acts.cpu().numpy().copy())
        for t in range(T):
            o, a, d, info = twin.step(acts[t])
            assert torch.equal(f.term_obs[t], o) and torch.equal(f.done[t], d) and torch.equal(f.obs[t + 1], info['next_obs']), (it, t)
    assert worker._graph is not None, worker.graph_error
    assert not np.array_equal(seen[1], seen[2])
    assert all(s.min() >= 0 and s.max() <= 3 for s in seen)
    env.close(); twin.close()
