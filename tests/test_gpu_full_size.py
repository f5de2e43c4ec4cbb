"""Full-size GPU checks at the shapes BASELINE.json names (C2 CIFAR-10, C3 hypersphere, C4 CelebA-64), where the CPU
oracle would take hours.  Size-independent properties stand in for it:

  * the tensor-core path against the exact-fp32 CUDA-core path (itself pinned to the oracle at small sizes) on a
    sample of the query rows, same tolerance contract as tests/test_gpu_kernels.py;
  * shard equivalence: statistics of the whole dataset == merge of the statistics of its row shards (what the
    multi-GPU path relies on), arg-min indices identical;
  * invariance under a permutation of the dataset rows (arg-min mapped through the permutation);
  * a noised copy of a training point at T -> 0 finds that point: arg-min = its index, logZ' = 0, <e> = 0.

Data are generated on the device (seeded) -- the shapes are 0.6 - 10 GB.
"""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def backend(cuda_device):
    from pdm_b200.backend import CudaBackend
    return CudaBackend(cuda_device)


def _make(kind, n, d, dev, seed):
    g = torch.Generator(device=dev).manual_seed(seed)
    if kind == "uniform":
        return torch.rand(n, d, device=dev, generator=g) * 2 - 1
    if kind == "sphere":                                         # utils/synthetic_datasets.py:14-17, radius sqrt(d)
        s = torch.randn(n, d, device=dev, generator=g)
        return s / (s.norm(dim=1, keepdim=True) / math.sqrt(d))
    if kind == "pixels":                                         # ToTensor + Normalize(0.5, 0.5) of uint8 pixels
        px = torch.randint(0, 256, (n, d), device=dev, generator=g, dtype=torch.uint8)
        return (px.float().div_(255.0, rounding_mode=None) - 0.5) / 0.5
    raise ValueError(kind)


def _queries(data, m, dev, seed):
    g = torch.Generator(device=dev).manual_seed(seed)
    idx = torch.randint(0, data.shape[0], (m,), device=dev, generator=g)
    temps = torch.logspace(-4, 4, m, device=dev)
    x = data[idx] + temps.sqrt()[:, None] * torch.randn(m, data.shape[1], device=dev, generator=g)
    return x, temps, idx


def _close(a, b, floor, rtol=1e-4, atol=2e-5, what=""):
    a, b = a.double(), b.double()
    tol = torch.maximum(rtol * b.abs() + atol, floor.double())
    bad = (a - b).abs() > tol
    assert not bad.any(), f"{what}: {int(bad.sum())} of {bad.numel()} outside tolerance, worst {(a - b).abs().max().item():.3e}"


CASES = [
    pytest.param("uniform", 50_000, 3072, 4096, id="C2-cifar10-shape"),
    pytest.param("pixels", 50_000, 3072, 4096, id="C2-8bit-pixels"),
    pytest.param("sphere", 100_000, 16384, 1024, id="C3-hypersphere"),
    pytest.param("uniform", 200_000, 12288, 1024, id="C4-celeba64-shape"),
]


@pytest.mark.parametrize("kind,n,d,m", CASES)
def test_full_size_properties(backend, kind, n, d, m):
    from pdm_b200 import EmpiricalDataset, PosteriorEngine, EngineConfig
    dev = backend.device
    data = _make(kind, n, d, dev, 100 + d)
    ds = EmpiricalDataset(data, backend=backend)
    eng = PosteriorEngine(ds, EngineConfig())
    assert eng.precision() == ("f16x2" if kind == "pixels" else "f16x3")
    x, temps, idx = _queries(data, m, dev, 7)
    st = eng.stats(x, temps)
    xn = (x.double() ** 2).sum(1)
    # a-priori round-off of an fp32 evaluation of |x|^2 - 2 x.y + |y|^2: 8 units of 2^-24 (|x|^2 + |y|^2) as in
    # tests/test_gpu_kernels.py, growing like sqrt(d) with the length of the fp32 accumulation (d/64 k-blocks here,
    # d terms in an SGEMM) -- 16 units at d = 16384
    units = max(8.0, math.sqrt(d) / 8)
    floor_e = units * 2.0 ** -24 * (xn + ds.y_norm.double().max()) / temps.double()
    keys = ("log_l", "mean_e", "entropy")

    # (1) tensor path vs the exact fp32 CUDA-core path on a sample of rows
    pick = torch.linspace(0, m - 1, 48, device=dev).long()
    ex = PosteriorEngine(ds, EngineConfig(precision="exact")).stats(x[pick], temps[pick])
    for k in keys:
        _close(st[k][pick], ex[k], (2 if k == "entropy" else 1) * floor_e[pick], what=f"{kind} tensor vs exact {k}")
    clear = (ex["log_l"] < 1e-3)                                  # one dominant neighbour: the arg-min is unambiguous
    assert torch.equal(st["argmin"][pick][clear], ex["argmin"][clear])
    # E_min against 1/2 ||x - y_argmin||^2 evaluated directly in fp64 (no norm expansion, no long fp32 sum); the exact
    # fp32 path accumulates d terms in one chain, so it is only held to sqrt(d)/8 times the floor
    e64 = 0.5 * ((x.double() - data[st["argmin"]].double()) ** 2).sum(1)
    floor_E = units * 2.0 ** -24 * (xn + ds.y_norm.double().max())
    _close(st["e_min"], e64, floor_E, atol=1e-5, what=f"{kind} e_min vs fp64")
    _close(ex["e_min"][clear], e64[pick][clear], floor_E[pick][clear] * max(1.0, math.sqrt(d) / 16), atol=1e-5,
           what=f"{kind} exact-path e_min vs fp64")

    # (2) the whole dataset == the merge of two row shards (the multi-GPU exchange step, on one device)
    half = n // 2
    parts = []
    inv_t = (1.0 / temps).contiguous()
    for lo, hi in ((0, half), (half, n)):
        shard = EmpiricalDataset(data[lo:hi], backend=backend, index_offset=lo, n_total=n, global_absmax=ds._absmax(),
                                 lattice_scale=ds.lattice_scale)
        e = PosteriorEngine(shard, EngineConfig())
        prep = e._prepare(x, m, None, None, None, e.precision(), False)
        parts.append(backend.reduce(e._local_partials(prep, m, inv_t, None, e.precision()), inv_t))
        del shard, e
    merged, argmin = backend.merge(torch.stack(parts), inv_t, n)
    from pdm_b200 import _cabi
    for k, row in (("log_l", _cabi.OUT_LOG_L), ("mean_e", _cabi.OUT_MEAN_E), ("entropy", _cabi.OUT_ENTROPY)):
        _close(merged[row], st[k], 0.25 * floor_e, rtol=2e-5, atol=2e-6, what=f"{kind} shard merge {k}")
    assert torch.equal(argmin, st["argmin"])

    # (3) T -> 0: a noised training point finds itself
    cold = temps < 1e-3
    assert cold.sum() > 0
    assert torch.equal(st["argmin"][cold], idx[cold])
    assert st["log_l"][cold].abs().max().item() < 1e-6 and st["mean_e"][cold].abs().max().item() < 1e-6
    del eng, ds, data
    torch.cuda.empty_cache()


def test_permutation_invariance(backend):
    from pdm_b200 import EmpiricalDataset, PosteriorEngine, EngineConfig
    dev = backend.device
    n, d, m = 50_000, 3072, 2048
    data = _make("uniform", n, d, dev, 5)
    x, temps, _ = _queries(data, m, dev, 9)
    st = PosteriorEngine(EmpiricalDataset(data, backend=backend), EngineConfig()).stats(x, temps)
    perm = torch.randperm(n, device=dev, generator=torch.Generator(device=dev).manual_seed(1))
    sp = PosteriorEngine(EmpiricalDataset(data[perm], backend=backend), EngineConfig()).stats(x, temps)
    xn = (x.double() ** 2).sum(1)
    floor_e = 8 * 2.0 ** -24 * (xn + (data.double() ** 2).sum(1).max()) / temps.double()
    for k in ("log_l", "mean_e", "entropy"):
        _close(sp[k], st[k], 0.25 * floor_e, rtol=2e-5, atol=2e-6, what=f"permutation {k}")
    assert torch.equal(sp["e_min"], st["e_min"])                  # the same dot products, in a different order of rows
    assert torch.equal(perm[sp["argmin"]], st["argmin"])
