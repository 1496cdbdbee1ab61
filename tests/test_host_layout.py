"""Host-side layout logic that needs no GPU: the pinned / device slabs of HotPathRunner slots and the flat
gradient / parameter layout of FlowTrainer."""
import torch

from fincflow_b200.stack import _carve, _carved_numel
from fincflow_b200.train import FlowTrainer, bump_versions


def test_carve_gives_aligned_disjoint_views():
    shapes = [(5, 12, 16, 16), (5,), (5, 8, 7, 7), (3,)]
    n = _carved_numel(shapes)
    flat = torch.zeros(n)
    views = _carve(flat, shapes)
    assert [tuple(v.shape) for v in views] == shapes
    offs = [(v.data_ptr() - flat.data_ptr()) // 4 for v in views]
    assert all(o % 64 == 0 for o in offs)                       # 256-byte boundaries
    for i, v in enumerate(views):
        v.fill_(i + 1)
    assert float(flat.sum()) == sum((i + 1) * v.numel() for i, v in enumerate(views))   # no overlap
    assert offs[-1] + views[-1].numel() <= n


def test_flow_trainer_flat_layout_on_cpu():
    m = torch.nn.Sequential(torch.nn.Linear(3, 5), torch.nn.Linear(5, 2))
    tr = FlowTrainer.__new__(FlowTrainer)          # layout only: no forward through a flow model
    tr.params = [p for p in m.parameters()]
    tr.world = 1
    tr._build_buckets(bucket_mb=1e-4)              # ~26 floats per bucket: several buckets
    offs = sorted(tr.offsets[p] for p in tr.params)
    assert all(o % 16 == 0 for o in offs)
    # reverse parameter order: the last layer's bias comes first
    assert tr.offsets[tr.params[-1]] == 0
    for p in tr.params:
        assert p.grad.data_ptr() == tr.flat_grad.data_ptr() + 4 * tr.offsets[p] and p.grad.shape == p.shape
    # buckets tile the buffer without gaps
    last = max(tr.offsets[p] + p.numel() for p in tr.params)     # only alignment padding may follow the last bucket
    assert tr.buckets[0][0] == 0 and last <= tr.buckets[-1][1] <= tr.flat_grad.numel()
    assert all(a[1] == b[0] for a, b in zip(tr.buckets, tr.buckets[1:]))
    assert sorted(id(p) for _, _, ms in tr.buckets for p in ms) == sorted(id(p) for p in tr.params)
    v = [p._version for p in tr.params]
    bump_versions(tr.params)
    assert [p._version for p in tr.params] == [x + 1 for x in v]
