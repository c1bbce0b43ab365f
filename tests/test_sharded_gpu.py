"""GPU (one device is enough): kronmult_batched_sharded_* with a single-rank communicator -- the redirect of the
shared output vectors into scratch and the final add run exactly as on N ranks, only the all-reduce is skipped --
plus the host-buffer entry points and the device-side ASGarD batch builder, all against the CPU oracle."""
import numpy as np
import pytest
import torch

from conftest import TOL
from kronmult993_b200 import batch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,d,dt", [(4, 5, torch.float64), (4, 5, torch.float32), (3, 3, torch.float64), (8, 4, torch.float64),
                                    (6, 6, torch.float64)])
def test_shared_outputs_single_rank(kron, oracle_mod, n, d, dt):
    nb = 300 if n ** d <= 5000 else 12
    hp = batch.make_problem(d, n, nb, dt, "cpu", seed=21, alias="ref", nb_distinct=4, lda=n + 2).to_host()
    expected = oracle_mod.run(hp, "oracle", threads=1)
    p = batch.from_host(hp, "cuda")
    A, i_, o_, w_ = p.pointer_arrays()
    comm = kron.Comm()
    N, s = hp.N, p.out_slab.element_size()
    keys = np.unique(hp.out_off)
    shared = [p.out_slab.data_ptr() + int(k) * s for k in keys[::2]]  # every other output goes through the scratch
    owner = np.zeros(len(shared), dtype=np.int32)
    st = torch.cuda.Stream()
    torch.cuda.synchronize()  # the pointer arrays were built on torch's default stream; `st` does not wait for it
    kron.kronmult_batched_sharded(d, n, A, p.lda, i_, o_, w_, p.nb, shared, comm, owner=owner, dtype=dt, stream=st)
    st.synchronize()
    err = oracle_mod.rel_l2(p.out_slab.cpu().numpy(), expected)
    assert err <= TOL[str(dt).split(".")[1]], err
    # not the owner: the vector keeps its values minus nothing -- its items' contributions are dropped by design
    comm.destroy()


def _host_arrays(hp):
    s = hp.in_slab.itemsize
    pa = (hp.mat_slab.ctypes.data + hp.mat_off * s).astype(np.uint64)
    pi = (hp.in_slab.ctypes.data + hp.in_off * s).astype(np.uint64)
    po = (hp.out_slab.ctypes.data + hp.out_off * s).astype(np.uint64)
    return pa, pi, po


@pytest.mark.parametrize("case", ["contiguous", "scattered", "overlapping", "cached_twice", "d0"])
def test_host_buffer_entry(kron, oracle_mod, case):
    """kronmult_batched_host_*: host pointer arrays of host vectors (signature of kronmult_omp/kronmult.hpp:77-80)."""
    dt = torch.float64
    if case == "d0":
        hp = batch.make_problem(0, 3, 9, dt, "cpu", seed=2, alias="runs", items_per_output=3).to_host()
    else:
        hp = batch.make_problem(3, 5, 700, dt, "cpu", seed=31, alias="runs", items_per_output=7, lda=6).to_host()
    N = hp.N
    if case == "scattered":
        groups = np.unique(hp.out_off)
        perm = np.random.default_rng(0).permutation(groups.size)
        spread = np.zeros(groups.size * 3 * N, dtype=hp.out_slab.dtype)  # every output vector 3N apart, shuffled
        lut = {int(g): int(perm[i]) * 3 * N for i, g in enumerate(groups)}
        for g in groups:
            spread[lut[int(g)]: lut[int(g)] + N] = hp.out_slab[g: g + N]
        hp.out_slab, hp.out_off = spread, np.array([lut[int(o)] for o in hp.out_off], dtype=np.int64)
    if case == "overlapping":
        # consecutive output vectors overlap by half: the reference's element-wise atomics make that legal
        # (kronmult.cu:126-129); the host path merges them into one span of the device copy
        hp.out_off = (hp.out_off // N) * (N // 2 + 1)
        hp.out_slab = hp.out_slab[: int(hp.out_off.max()) + N].copy()
    expected = oracle_mod.run(hp, "oracle", threads=1)
    pa, pi, po = _host_arrays(hp)
    before = torch.cuda.current_device()
    reps = 2 if case == "cached_twice" else 1
    for _ in range(reps):
        kron.kronmult_batched_host(hp.d, hp.n, pa.ctypes.data if hp.d else 0, hp.lda, pi.ctypes.data, po.ctypes.data, 0,
                                   hp.nb, dtype=dt, device=0)
    assert torch.cuda.current_device() == before
    if case == "cached_twice":
        hp2 = batch.make_problem(3, 5, 700, dt, "cpu", seed=31, alias="runs", items_per_output=7, lda=6).to_host()
        once = oracle_mod.run(hp2, "oracle", threads=1)
        hp2.out_slab[:] = once
        expected = oracle_mod.run(hp2, "oracle", threads=1)
    err = oracle_mod.rel_l2(hp.out_slab, expected)
    assert err <= 1e-12, err


def test_host_entry_on_every_device_of_one_process(kron, oracle_mod):
    """The shared-memory opt-in and the occupancy of a kernel are per device: one host thread drives device 0, then
    device 1, through kernels that need more than 48 KiB (n=4 d=5, DMMA n=8 d=4, a 64 KiB pairtile shape)."""
    ndev = torch.cuda.device_count()
    if ndev < 2:
        pytest.skip("needs >= 2 GPUs")
    for (d, n, nb) in [(5, 4, 200), (4, 8, 40), (5, 6, 12)]:
        for dev in range(min(ndev, 3)):
            hp = batch.make_problem(d, n, nb, torch.float64, "cpu", seed=dev + d, alias="runs", items_per_output=5).to_host()
            expected = oracle_mod.run(hp, "oracle", threads=1)
            pa, pi, po = _host_arrays(hp)
            kron.kronmult_batched_host(d, n, pa.ctypes.data, hp.lda, pi.ctypes.data, po.ctypes.data, 0, hp.nb,
                                       dtype=torch.float64, device=dev)
            err = oracle_mod.rel_l2(hp.out_slab, expected)
            assert err <= 1e-12, (d, n, dev, err)
    assert torch.cuda.current_device() == 0


@pytest.mark.parametrize("n,d,dt", [(4, 5, torch.float64), (3, 2, torch.float64), (5, 3, torch.float32), (6, 6, torch.float64)])
def test_device_side_asgard_batch_builder(kron, oracle_mod, n, d, dt):
    """kronmult_build_batch_* + kronmult_batched_const_*: items (row element, column element, term), windows into
    d x nterms one-dimensional coefficient matrices, x shared by all rows -- against the oracle on the same batch
    written out on the host."""
    gen = torch.Generator().manual_seed(7)
    nterms, ncell = 3, 4
    nelem = 9 if n ** d <= 5000 else 3
    lda = ncell * n
    N = n ** d
    cells = torch.randint(0, ncell, (nelem, d), generator=gen, dtype=torch.int32)
    coeff = torch.randn(nterms * d, lda * lda, generator=gen, dtype=torch.float64).to(dt)
    x = torch.randn(nelem * N, generator=gen, dtype=torch.float64).to(dt)
    y = torch.randn(nelem * N, generator=gen, dtype=torch.float64).to(dt)
    rows, cols = (1, nelem), (0, nelem - 1)
    nb = (rows[1] - rows[0]) * (cols[1] - cols[0]) * nterms
    # the same batch on the host, for the oracle
    mat_off, in_off, out_off = [], [], []
    for i in range(*rows):
        for j in range(*cols):
            for t in range(nterms):
                in_off.append(j * N)
                out_off.append(i * N)
                for dim in range(d):
                    mat_off.append((t * d + dim) * lda * lda + n * int(cells[i, dim]) + n * int(cells[j, dim]) * lda)
    # the oracle clobbers its inputs like the reference (kronmult_omp/kronmult.hpp:47-51): give every item a private
    # copy of its column element's vector (what an ASGarD caller has to do without the const entry points)
    xn = x.numpy()
    hp = batch.HostProblem(d, n, lda, nb, coeff.numpy().ravel().copy(), np.array(mat_off, dtype=np.int64),
                           np.concatenate([xn[o:o + N] for o in in_off]), np.arange(nb, dtype=np.int64) * N,
                           y.numpy().copy(), np.array(out_off, dtype=np.int64))
    expected = oracle_mod.run(hp, "oracle", threads=1)
    dev = "cuda"
    cx, cc, xx, yy = cells.to(dev), coeff.to(dev), x.to(dev), y.to(dev)
    cptr = (cc.data_ptr() + torch.arange(nterms * d, dtype=torch.int64) * (lda * lda * cc.element_size())).to(dev)
    A, i_, o_, got = kron.build_asgard_batch(d, n, lda, cx, cptr, nterms, rows, cols, xx, yy)
    assert got == nb
    s = xx.element_size()
    assert torch.equal(i_.cpu(), torch.tensor(in_off, dtype=torch.int64) * s + xx.data_ptr())
    assert torch.equal(o_.cpu(), torch.tensor(out_off, dtype=torch.int64) * s + yy.data_ptr())
    assert torch.equal(A.cpu(), torch.tensor(mat_off, dtype=torch.int64) * s + cc.data_ptr())
    ws = None
    if kron.needs_workspace(d, n, dt):
        wslab = torch.zeros(nb * N, dtype=dt, device=dev)
        ws = (wslab.data_ptr() + torch.arange(nb, dtype=torch.int64) * (N * s)).to(dev)
    kron.kronmult_batched_const(d, n, A, lda, i_, o_, ws, nb, dtype=dt)
    err = oracle_mod.rel_l2(yy.cpu().numpy(), expected)
    assert err <= TOL[str(dt).split(".")[1]], err
    assert torch.equal(xx.cpu(), x)  # the shared inputs are untouched
