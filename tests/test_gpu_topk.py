"""tcgen05 top-K scoring kernel vs fp32 PyTorch (TF32 tolerance) and brute-force top-K."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    torch.cuda.set_device(0)
    return torch.device("cuda", 0)


@pytest.mark.parametrize("nq,ni,k", [(128, 256, 64), (200, 1000, 64), (50, 333, 16), (300, 5000, 128), (130, 700, 10)])
def test_scores_match_fp32_matmul(dev, nq, ni, k):
    from fps_b200.models.mf.device_topk import DeviceTopK

    stride = (k + 3) // 4 * 4
    g = torch.Generator(device="cpu").manual_seed(nq + ni)
    items = torch.zeros(ni, stride); items[:, :k] = torch.randn(ni, k, generator=g)
    q = torch.zeros(nq, stride); q[:, :k] = torch.randn(nq, k, generator=g)
    items, q = items.to(dev), q.to(dev)
    got = DeviceTopK(items).scores(q_local=q)
    ref = q.double() @ items.double().T
    err = (got.double() - ref).abs().max().item()
    assert err < 2e-2 * (k ** 0.5), err          # TF32: 10-bit mantissa inputs, fp32 accumulate
    assert torch.allclose(got.double(), ref, rtol=5e-3, atol=5e-2)


def test_topk_equals_brute_force_with_pull_from_ps(dev):
    from fps_b200.models.mf.device_topk import DeviceTopK
    from fps_b200.store.sharded_table import ShardedTable

    k, nu, ni, K = 64, 3000, 20000, 100
    users = ShardedTable(nu, k, seed=1, init_range=(-1.0, 1.0))
    g = torch.Generator(device="cpu").manual_seed(0)
    scale = torch.rand(ni, 1, generator=g) * 2 + 0.1
    items = (torch.randn(ni, k, generator=g) * scale).to(dev).contiguous()
    q_ids = torch.randint(0, nu, (517,), generator=g).to(dev)
    tk = DeviceTopK(items)
    sc, rows = tk.topk(K, q_ids=q_ids, q_table=users)
    full = tk.scores(q_ids=q_ids, q_table=users)            # same TF32 arithmetic
    ref = torch.topk(full, K, dim=1)
    assert torch.equal(sc, ref.values)
    # item sets must agree wherever scores are distinct
    same = (rows == ref.indices) | (sc == torch.roll(sc, 1, 1)) | (sc == torch.roll(sc, -1, 1))
    assert same.all()
    # and against exact fp32: the exact top-10 must be inside the TF32 top-100
    u = users.pull(q_ids)
    exact = torch.topk(u @ items.T, 10, dim=1).indices
    hit = (exact[:, :, None] == rows[:, None, :]).any(-1).float().mean().item()
    assert hit > 0.999, hit
    sc2, rows2 = tk.topk(10, q_ids=q_ids, q_table=users, rescore=True)
    assert (sc2[:, :-1] >= sc2[:, 1:]).all()
    users.close()


def test_topk_small_item_table_and_k_larger_than_tiles(dev):
    from fps_b200.models.mf.device_topk import DeviceTopK, merge_partial_topk

    items = torch.randn(300, 32, device=dev)
    q = torch.randn(40, 32, device=dev)
    tk = DeviceTopK(items)
    sc, rows = tk.topk(50, q_local=q)
    ref = torch.topk(tk.scores(q_local=q), 50, dim=1)
    assert torch.equal(sc, ref.values)
    ms, mi = merge_partial_topk(torch.cat([sc, sc - 1000], 1), torch.cat([rows, rows], 1), 50)
    assert torch.equal(ms, sc) and torch.equal(mi, rows)
