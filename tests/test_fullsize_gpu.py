"""GPU: BASELINE.json configurations at FULL size.

Every output element is compared with a plain PyTorch (einsum / index_add_) evaluation of the same
operator on the same device data (tests/_torch_ref.py), and a strided subset of output groups is
additionally recomputed by the CPU oracle from exactly the items that feed them.
Sizes: C2 16 Mi items, C3 1 Mi x 4096, C4 512 Ki x 4096 (aliased and distinct), C5 8 Mi x 1024.
"""
import numpy as np
import pytest
import torch

import _torch_ref as tref
from conftest import TOL
from kronmult993_b200 import batch

pytestmark = pytest.mark.gpu

CONFIGS = {
    # name: (d, n, nb, dtype, items_per_output)
    "c1": (3, 4, 65536, torch.float64, 1),
    "c2": (2, 2, 1 << 24, torch.float64, 1),
    "c3": (6, 4, 1 << 20, torch.float64, 32),
    "c4a": (4, 8, 1 << 19, torch.float64, 1),
    "c4b": (4, 8, 1 << 19, torch.float64, 32),
    "c5_f64": (5, 4, 1 << 23, torch.float64, 32),
    "c5_f32": (5, 4, 1 << 23, torch.float32, 32),
}


@pytest.mark.parametrize("name", list(CONFIGS))
def test_baseline_config_full_size(kron, oracle_mod, name):
    d, n, nb, dt, r = CONFIGS[name]
    need = nb * n ** d * (8 if dt == torch.float64 else 4) * 1.35 + (6 << 30)
    free, _ = torch.cuda.mem_get_info()
    if free < need:
        pytest.skip(f"{name} needs {need / 2**30:.0f} GiB of device memory")
    p = batch.make_problem(d, n, nb, dt, "cuda", seed=993, alias="runs" if r > 1 else "distinct",
                           items_per_output=r)
    tol = TOL["float64" if dt == torch.float64 else "float32"]
    expected = tref.reference_output(p)
    # oracle on a strided subset of output groups, from the pristine data
    groups = torch.arange(0, p.n_outputs, max(1, p.n_outputs // 48), device="cuda")[:48]
    hp, groups = p.select_outputs_to_host(groups)
    exp_sub = oracle_mod.run(hp, "oracle", threads=None)

    kron.run_problem(p)
    torch.cuda.synchronize()
    err = tref.rel_l2(p.out_slab, expected)
    assert err <= tol, f"{name}: rel-L2 vs torch reference {err:.3e} ({kron.last_path()})"
    N = p.N
    got_sub = p.out_slab.view(-1, N)[groups].flatten().cpu().numpy()
    err2 = oracle_mod.rel_l2(got_sub, exp_sub)
    assert err2 <= tol, f"{name}: rel-L2 vs CPU oracle on {groups.numel()} groups {err2:.3e}"
    del expected, p
    torch.cuda.empty_cache()
