"""GPU tests of the edge builders against a brute-force PyTorch restatement of radius_graph / knn_graph
semantics (same-graph pairs, loop flag, neighbour cap in index order / k nearest with index tie-break)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _mols(sizes, seed, scale=2.0):
    g = torch.Generator().manual_seed(seed)
    pos = scale * torch.randn(sum(sizes), 3, generator=g)
    batch = torch.cat([torch.full((n,), i, dtype=torch.long) for i, n in enumerate(sizes)])
    return pos, batch


def _ref_radius(pos, batch, r, loop, max_nb):
    same = batch[:, None] == batch[None, :]
    dx = pos[:, None, :] - pos[None, :, :]
    ok = ((dx * dx).sum(-1) < r * r) & same   # squared fp32 distances, like the kernel
    if not loop:
        ok &= ~torch.eye(len(pos), dtype=torch.bool)
    src, dst = [], []
    for i in range(len(pos)):
        js = torch.nonzero(ok[i]).flatten()[:max_nb]
        src += js.tolist(); dst += [i] * len(js)
    return torch.tensor([src, dst], dtype=torch.long)


def _ref_knn(pos, batch, k, loop):
    dx = pos[:, None, :] - pos[None, :, :]
    d = (dx * dx).sum(-1)
    same = batch[:, None] == batch[None, :]
    d = torch.where(same, d, torch.full_like(d, float("inf")))
    if not loop:
        d.fill_diagonal_(float("inf"))
    src, dst = [], []
    for i in range(len(pos)):
        order = torch.sort(d[i], stable=True).indices
        js = [int(j) for j in order[:k] if torch.isfinite(d[i, j])]
        src += js; dst += [i] * len(js)
    return torch.tensor([src, dst], dtype=torch.long)


@pytest.mark.parametrize("loop", [False, True])
@pytest.mark.parametrize("r,max_nb", [(1.5, 32), (3.0, 32), (6.0, 5)])
def test_radius_graph_matches_bruteforce(r, max_nb, loop):
    from moldiff_b200.graph_build import radius_graph
    pos, batch = _mols([7, 1, 24, 39, 2, 64], seed=1)
    ei = radius_graph(pos.cuda(), r, batch=batch.cuda(), loop=loop, max_num_neighbors=max_nb).cpu()
    ref = _ref_radius(pos, batch, r, loop, max_nb)
    assert ei.dtype == torch.int64 and torch.equal(ei, ref)
    assert bool((batch[ei[0]] == batch[ei[1]]).all())


@pytest.mark.parametrize("k", [1, 4, 32])
@pytest.mark.parametrize("flow", ["source_to_target", "target_to_source"])
def test_knn_graph_matches_bruteforce(k, flow):
    from moldiff_b200.graph_build import knn_graph
    pos, batch = _mols([7, 1, 24, 39, 2, 64], seed=2)
    ei = knn_graph(pos.cuda(), k, batch=batch.cuda(), flow=flow).cpu()
    ref = _ref_knn(pos, batch, k, loop=False)
    if flow == "target_to_source":
        ref = ref.flip(0)
    assert torch.equal(ei, ref)


def test_builders_edge_cases_and_errors():
    from moldiff_b200 import engine
    from moldiff_b200.graph_build import knn_graph, radius_graph
    empty = torch.zeros(0, 3, device="cuda")
    assert radius_graph(empty, 1.0).shape == (2, 0) and knn_graph(empty, 3).shape == (2, 0)
    one = torch.zeros(1, 3, device="cuda")
    assert radius_graph(one, 1.0).shape == (2, 0) and radius_graph(one, 1.0, loop=True).tolist() == [[0], [0]]
    pos, _ = _mols([10], seed=3)
    full = radius_graph(pos.cuda(), 1e3, batch=None)                    # no batch = one graph: complete graph without loops
    assert full.shape[1] == 90
    with pytest.raises(engine.MoldiffB200Error):
        radius_graph(pos, 1.0)                                          # CPU tensor: no CPU path
    with pytest.raises(engine.MoldiffB200Error):
        knn_graph(pos.cuda(), 33)                                       # k > 32
    with pytest.raises(engine.MoldiffB200Error):
        radius_graph(pos.cuda(), 1.0, batch=torch.tensor([1, 0] * 5).cuda())   # unsorted batch


def test_radius_edges_feed_the_denoiser_network():
    """A cutoff-sparsified edge list from radius_graph runs through NodeEdgeNet (any edge_index is accepted)."""
    from moldiff_b200.graph_build import radius_graph
    from moldiff_b200.nets import NodeEdgeNet
    pos, batch = _mols([12, 30, 20], seed=4)
    dev = torch.device("cuda:0")
    ei = radius_graph(pos.to(dev), 3.0, batch=batch.to(dev))
    torch.manual_seed(0)
    net = NodeEdgeNet(256, 64, num_blocks=2, cutoff=15.0, use_gate=True).to(dev).eval()
    N, E = len(pos), ei.shape[1]
    g = torch.Generator().manual_seed(5)
    h_node, h_edge = torch.randn(N, 256, generator=g).to(dev), torch.randn(E, 64, generator=g).to(dev)
    with torch.no_grad():
        out = net(h_node, pos.to(dev), h_edge, ei, torch.rand(N, 1, generator=g).to(dev), torch.rand(E, 1, generator=g).to(dev))
    assert all(torch.isfinite(o).all() for o in out) and out[2].shape == (E, 64)
