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
    model = DeviceWideAndDeep(slots, fields, emb_dim=8, hidden=32, learning_rate=0.05, pull_limit=64, seed=1)
    losses = []
    for _ in range(150):
        ids = torch.randint(0, 2000, (B, fields), generator=g)
        y = (w_true[ids].sum(1) > 0).float()
        losses.append(model.step(ids.to(dev), y.to(dev)))
    assert sum(losses[-10:]) / 10 < 0.8 * sum(losses[:10]) / 10
    ids = torch.randint(0, 2000, (2000, fields), generator=g)
    acc = ((model.predict(ids.to(dev)).cpu() > 0.5).float() == (w_true[ids].sum(1) > 0).float()).float().mean()
    assert acc > 0.7, acc
    model.close()
