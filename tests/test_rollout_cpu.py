"""CPU: host-side logic of the rollout module (policy shapes / determinism, GAE against a scalar
restatement).  The collection loop itself needs the CUDA engine: tests/test_gpu_rollout.py."""
import numpy as np
import torch

from paintrl_b200.rollout import MlpPolicy, RolloutFragment, gae


def test_policy_shapes_and_determinism():
    p1 = MlpPolicy(6, 4, device=torch.device('cpu'), seed=3)
    p2 = MlpPolicy(6, 4, device=torch.device('cpu'), seed=3)
    obs = torch.rand(17, 6, dtype=torch.float64)
    a1, l1, v1 = p1.act(obs)
    a2, l2, v2 = p2.act(obs)
    assert a1.shape == (17,) and a1.dtype == torch.int64 and l1.shape == (17,) and v1.shape == (17,)
    assert torch.equal(a1, a2) and torch.equal(l1, l2) and torch.equal(v1, v2)
    assert int(a1.min()) >= 0 and int(a1.max()) < 4 and bool((l1 <= 0).all())
    assert [tuple(w.shape) for w, _ in p1.layers] == [(6, 256), (256, 128)]       # paint_ppo.py:180
    logits, _ = p1.forward(obs)
    assert torch.allclose(torch.log_softmax(logits, 1).gather(1, a1[:, None])[:, 0], l1)
    pc = MlpPolicy(16, 2, device=torch.device('cpu'), seed=1, discrete=False)
    a, logp, v = pc.act(torch.rand(5, 16))
    assert a.shape == (5, 2) and a.dtype == torch.float64 and logp.shape == (5,)


def test_gae_matches_scalar_restatement():
    T, B = 9, 5
    g = torch.Generator().manual_seed(0)
    f = RolloutFragment(T, B, 3, (), torch.device('cpu'))
    f.actual.copy_(torch.rand(T, B, generator=g, dtype=torch.float64))
    f.value.copy_(torch.rand(T + 1, B, generator=g))
    f.done.copy_((torch.rand(T, B, generator=g) < 0.25).to(torch.uint8))
    adv, target = gae(f, gamma=0.9, lam=0.8)
    ref = np.zeros((T, B), dtype=np.float32)
    for b in range(B):
        last = np.float32(0)
        for t in range(T - 1, -1, -1):
            nd = np.float32(1 - int(f.done[t, b]))
            delta = np.float32(f.actual[t, b]) + np.float32(0.9) * f.value[t + 1, b].numpy() * nd - f.value[t, b].numpy()
            last = delta + np.float32(0.9 * 0.8) * nd * last
            ref[t, b] = last
    assert np.allclose(adv.numpy(), ref, rtol=1e-5, atol=1e-6)
    assert np.allclose(target.numpy(), ref + f.value[:T].numpy(), rtol=1e-5, atol=1e-6)
