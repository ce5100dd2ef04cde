"""Host logic of the owner-row strip engine, on CPU: the planner's tasks must tile every shard's slice of
the packed array exactly once (the reference writes every element of XX once, int2e.f90:290-307) and must
store exactly the integrals whose two shell pairs pass the reference's EIJ*EGH >= 1e-14 rule (:257).
`myqc_eri_plan_check` replays what the kernels write with the kernels' own index arithmetic
(csrc/strip_geom.hpp); no device is involved."""
import numpy as np
import pytest

import myqc_b200 as Q
from conftest import product_system


@pytest.mark.parametrize("name", ["H", "H2", "HeH", "Be", "O_singlet", "HF", "OH", "CO", "NO", "CO2", "h2o_2", "c4h10", "h2o_8"])
def test_tasks_tile_the_packed_array_exactly_once(name, tmp_path, capfd):
    s = product_system(name, tmp_path)
    r = Q.plan_check(s)
    assert r["errors"] == 0
    assert r["slice_elems"] == s.nunique == r["zero_filled"]
    assert r["stored"] == r["expected"] > 0


@pytest.mark.parametrize("name,nsh", [("CO2", 2), ("CO2", 3), ("CO2", 8), ("h2o_4", 3), ("h2o_4", 8), ("h2o_8", 4), ("c4h10", 5)])
def test_shards_tile_the_array_and_each_shard_is_complete(name, nsh, tmp_path, capfd):
    """Shards own contiguous row blocks; a shard's tasks depend only on the rows it owns, so the shards compute
    what the unsharded plan computes.  More shards than cut points leave empty shards (no tasks, no elements)."""
    s = product_system(name, tmp_path)
    off = Q.shard_layout(s, nsh)
    assert off[0] == 0 and off[-1] == s.nunique and np.all(np.diff(off) >= 0)
    whole = Q.plan_check(s)
    stored = 0
    for k in range(nsh):
        r = Q.plan_check(s, k, nsh)
        assert r["errors"] == 0
        assert r["slice_elems"] == off[k + 1] - off[k] == r["zero_filled"]
        assert r["stored"] == r["expected"]
        if r["slice_elems"] == 0:
            assert r["tasks"] == 0
        stored += r["stored"]
    assert stored == whole["stored"]


def test_h2o16_baseline_config_plan(tmp_path, capfd):
    """BASELINE.json configs[2] (112 functions, 20 024 956 unique integrals): whole and as 8 shards."""
    s = product_system("h2o_16", tmp_path)
    r = Q.plan_check(s)
    assert r["errors"] == 0 and r["slice_elems"] == 20024956 and r["stored"] == r["expected"]
    sizes = []
    for k in range(8):
        rk = Q.plan_check(s, k, 8)
        assert rk["errors"] == 0 and rk["stored"] == rk["expected"]
        sizes.append(rk["slice_elems"])
    assert sum(sizes) == 20024956 and min(sizes) > 0
