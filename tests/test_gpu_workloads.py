"""North-star workloads beyond the reference's suite: skip-gram (dim=300) and wide-&-deep CTR."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_skipgram_step_matches_fp32_reference_and_learns():
    from fps_b200.models.w2v import DeviceSkipGram

    torch.cuda.set_device(0)
    dev = torch.device("cuda", 0)
    V, D = 2000, 300
    m = DeviceSkipGram(V, D, learning_rate=0.05, negative=0, seed=1)
    m.w_out.local.uniform_(-0.05, 0.05)
    Win, Wout = m.w_in.local[:, :D].clone(), m.w_out.local[:, :D].clone()
    c = torch.randperm(V, device=dev)[:500].int(); o = torch.randperm(V, device=dev)[:500].int()
    m.step(c, o)
    torch.cuda.synchronize()
    u, v = Win[c.long()], Wout[o.long()]
    g = (0.05 * (1 - torch.sigmoid((u * v).sum(1))))[:, None]
    torch.testing.assert_close(m.w_in.local[:, :D], Win.clone().index_add_(0, c.long(), g * v), rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(m.w_out.local[:, :D], Wout.clone().index_add_(0, o.long(), g * u), rtol=1e-5, atol=1e-6)
    # tiny corpus: words 2k and 2k+1 always co-occur -> their vectors must become similar
    m2 = DeviceSkipGram(64, 32, learning_rate=0.1, negative=4, seed=2)
    a = torch.arange(0, 64, 2, device=dev).int().repeat(8); b = a + 1
    for _ in range(300):
        m2.step(a, b); m2.step(b, a)
    m2.check_finite()
    partner = m2.score(torch.arange(0, 64, 2, device=dev), torch.arange(1, 64, 2, device=dev)).mean()
    stranger = m2.score(torch.arange(0, 62, 2, device=dev), torch.arange(3, 64, 2, device=dev)).mean()
    assert partner > 0.8 and partner > stranger + 0.3, (partner.item(), stranger.item())
    assert m2.similarity(torch.arange(4, device=dev), torch.arange(4, device=dev)).min() > 0.999
    m.close(); m2.close()


def test_wide_and_deep_ctr_learns_with_pull_limit_64():
    from fps_b200.models.ctr import DeviceWideAndDeep

    torch.cuda.set_device(0)
    dev = torch.device("cuda", 0)
    g = torch.Generator(device="cpu").manual_seed(0)
    slots, fields, B = 50_000, 6, 512
    w_true = torch.randn(slots, generator=g)
    model = DeviceWideAndDeep(slots, fields, emb_dim=8, learning_rate=0.05, pull_limit=64, seed=1)
    losses = []
    for _ in range(150):
        ids = torch.randint(0, 2000, (B, fields), generator=g)
        y = (w_true[ids].sum(1) > 0).float()
        losses.append(model.step(ids.to(dev), y.to(dev)).clone())          # device tensors: no sync per step
    losses = [float(l[0] / l[1]) for l in losses]
    assert sum(losses[-10:]) / 10 < 0.8 * sum(losses[:10]) / 10
    ids = torch.randint(0, 2000, (2000, fields), generator=g)
    acc = ((model.predict(ids.to(dev)).cpu() > 0.5).float() == (w_true[ids].sum(1) > 0).float()).float().mean()
    assert acc > 0.7, acc
    assert model.credit_stalls() >= 0 and model.table._credits(64, dev)[0].item() == 64   # every credit returned
    model.close()


def test_ctr_fused_tower_matches_autograd():
    """One step of the fused tower kernel (forward, BCE, backward, dense SGD, row gradients) against the same
    model written with torch autograd in fp32."""
    from fps_b200.ops import native

    torch.cuda.set_device(0)
    dev = torch.device("cuda", 0)
    g = torch.Generator(device="cpu").manual_seed(3)
    F, E, H, B, stride, lr = 26, 8, 256, 777, 12, 0.05
    rows = (torch.randn(B * F, stride, generator=g) * 0.3).to(dev)
    rows[:, E + 1:] = 0
    y = (torch.rand(B, generator=g) < 0.4).float().to(dev)
    W1 = (torch.randn(F * E, H, generator=g) * 0.1).to(dev)
    w = {"W1": W1.clone(), "W1T": W1.t().contiguous(), "b1": (torch.randn(H, generator=g) * 0.1).to(dev),
         "w2": (torch.randn(H, generator=g) * 0.1).to(dev), "b2": (torch.randn(1, generator=g) * 0.1).to(dev)}
    ref = {k: v.clone().requires_grad_(True) for k, v in w.items() if k != "W1T"}
    r = rows.clone().requires_grad_(True)
    x = r[:, :E].reshape(B, F * E)
    logit = torch.relu(x @ ref["W1"] + ref["b1"]) @ ref["w2"] + ref["b2"] + r[:, E].reshape(B, F).sum(1)
    loss_ref = torch.nn.functional.binary_cross_entropy_with_logits(logit, y, reduction="sum")
    loss_ref.backward()
    grads = {k: torch.zeros_like(v) for k, v in w.items() if k != "W1T"}
    d_rows = torch.empty_like(rows)
    loss = torch.zeros(2, device=dev)
    prob = torch.empty(B, device=dev)
    native.ctr_step(rows, y, F, E, w, grads, d_rows, loss, lr, train=True, prob=prob)
    torch.cuda.synchronize()
    torch.testing.assert_close(loss[0], loss_ref.detach(), rtol=1e-4, atol=1e-3)
    assert loss[1].item() == B
    torch.testing.assert_close(prob, torch.sigmoid(logit.detach()), rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(d_rows, -lr * r.grad, rtol=1e-3, atol=1e-6)
    for k in ("W1", "b1", "w2", "b2"):                       # dense SGD with the MEAN gradient
        torch.testing.assert_close(w[k], ref[k].detach() - lr / B * ref[k].grad, rtol=1e-4, atol=1e-6)
        assert grads[k].abs().max().item() == 0.0            # zeroed for the next step
    torch.testing.assert_close(w["W1T"], w["W1"].t())
