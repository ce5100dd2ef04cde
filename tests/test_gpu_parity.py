"""Parity tests proper (need a B200): the CUDA path through the C-ABI against the CPU oracle and
the committed golden fixtures.  Tolerance: BASELINE.json's north star -- 1e-10 Hartree absolute per
integral."""
import os

import numpy as np
import pytest

import myqc_b200 as Q
from conftest import EXAMPLES, GOLDEN, INPUTS, example_zmat, oracle_system, product_system
from oracle import oracle as O

pytestmark = pytest.mark.gpu
TOL = 1.0e-10  # Hartree, absolute, per integral (BASELINE.json north_star)


@pytest.fixture(scope="module", autouse=True)
def _need_gpu():
    if Q.device_count() < 1:
        pytest.fail("GPU tests need a CUDA device; the ERI engine has no CPU fallback")


@pytest.mark.parametrize("name", EXAMPLES)
def test_examples_against_golden_and_oracle(name, tmp_path, oracle_inputs):
    s = product_system(name, tmp_path)
    got = Q.eri_packed(s)
    gold = np.load(os.path.join(GOLDEN, f"packed_{name}.npy"))
    assert got.shape == gold.shape
    assert np.abs(got - gold).max() < TOL
    mol, b, ft = oracle_system(name, oracle_inputs)
    assert np.abs(got - O.int2e_packed(mol, b, ft)).max() < TOL
    # what the reference's screen leaves exactly zero stays (numerically) zero
    assert np.abs(got[gold == 0.0]).max(initial=0.0) < 1e-14


@pytest.mark.parametrize("name", ["CO2", "NO", "h2o_2"])
def test_dense_xx_all_eight_images(name, tmp_path, oracle_inputs):
    """proc2e mirror: dense XX(i,j,g,h) with fillsym's 8 images (int2e.f90:290-304,540-554)."""
    s = product_system(name, tmp_path)
    xx = Q.proc2e(s.bas, s.basinfo, s.atoms, s.options, 1000.0, s.nnuc, s.xyz, s.set, s.setinfo, s.maxL, s.ftab)
    mol, b, ft = oracle_system(name, oracle_inputs)
    ref, _ = O.int2e_dense(mol, b, ft)
    assert xx.shape == ref.shape and np.abs(xx - ref).max() < TOL
    for perm in [(1, 0, 2, 3), (0, 1, 3, 2), (1, 0, 3, 2), (2, 3, 0, 1), (3, 2, 0, 1), (2, 3, 1, 0), (3, 2, 1, 0)]:
        assert np.array_equal(xx, xx.transpose(perm))


@pytest.mark.parametrize("name", ["CO2", "OH"])
def test_int2e_drop_in_writes_the_reference_xx_record(name, tmp_path, oracle_inputs):
    """PROGRAM int2e in a job directory: same inputs, same XX record (consumers do READ(9) XX)."""
    s = product_system(name, tmp_path)
    assert Q.int2e_main(str(tmp_path), 1) == 0
    assert not (tmp_path / "error").exists()
    raw = open(tmp_path / "XX", "rb").read()
    n = s.norb
    assert len(raw) == 8 * n ** 4 + 8
    xx = np.frombuffer(raw[4:-4], dtype="<f8").reshape((n,) * 4, order="F")
    mol, b, ft = oracle_system(name, oracle_inputs)
    ref, _ = O.int2e_dense(mol, b, ft)
    assert np.abs(xx - ref).max() < TOL
    # basinfo was (re)written for the downstream stages; fmem is net unchanged and readable
    assert int(open(tmp_path / "basinfo").read().split()[1]) == n
    assert abs(float(open(tmp_path / "fmem").read().split()[0]) - 1000.0) < 1e-6
    # the SCF energy the consumers would get from this XX
    S, H = O.int1e(mol, b, ft)
    nA, nB = O.electrons(mol)
    if nA == nB:
        E, _, _ = O.scf_rhf(S, H, np.array(xx), nA + nB, O.nuclear_repulsion(mol))
        assert abs(E - (-183.32315970625)) < 1e-9  # 1e-9 Eh, north star


def test_int2e_executable_found_on_path_like_the_driver_does(tmp_path, oracle_inputs):
    """The reference driver runs `int2e` by name in the job directory (CALL EXECUTE_COMMAND_LINE('int2e'),
    src/myQC/myQC.f90:54) and then only looks for the `error` file (:55-59).  Same here: PATH lookup, cwd = job
    directory, no arguments; a second run finds XX and leaves it alone (int2e.f90:58-63)."""
    import subprocess
    s = product_system("HF", tmp_path)
    exe_dir = os.path.join(os.path.dirname(Q.__file__), "csrc")
    assert os.path.exists(os.path.join(exe_dir, "int2e"))
    env = dict(os.environ, PATH=exe_dir + os.pathsep + os.environ.get("PATH", ""))
    out = subprocess.run("int2e", shell=True, cwd=str(tmp_path), env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and not (tmp_path / "error").exists(), out.stdout + out.stderr
    n = s.norb
    xx = Q.read_xx(str(tmp_path / "XX"), n)
    mol, b, ft = oracle_system("HF", oracle_inputs)
    ref, _ = O.int2e_dense(mol, b, ft)
    assert np.abs(xx - ref).max() < TOL
    stamp = os.stat(tmp_path / "XX").st_mtime_ns
    again = subprocess.run("int2e", shell=True, cwd=str(tmp_path), env=env, capture_output=True, text=True, timeout=300)
    assert again.returncode == 0 and os.stat(tmp_path / "XX").st_mtime_ns == stamp
    # failure path: without Ftab the program touches `error`, as the reference does (int2e.f90:174-178)
    os.remove(tmp_path / "XX")
    os.remove(tmp_path / "Ftab")
    subprocess.run("int2e", shell=True, cwd=str(tmp_path), env=env, capture_output=True, text=True, timeout=300)
    assert (tmp_path / "error").exists() and not (tmp_path / "XX").exists()


@pytest.mark.parametrize("name,ngpu", [("CO2", 2), ("h2o_4", 3)])
def test_dense_xx_assembled_in_slabs_matches_single_device(name, ngpu, tmp_path, monkeypatch):
    """myqc_eri_dense(ngpu > 1): shards of the packed array, then one slab XX(:,:,:,h0:h1) per device.  With fewer
    devices than shards MYQC_OVERSUBSCRIBE=1 maps them round-robin onto the devices there are (test hook), so the
    in-process multi-device path runs on a one-GPU box too; on a multi-GPU box it uses distinct devices."""
    s = product_system(name, tmp_path)
    one = Q.eri_dense(s, ngpu=1)
    if Q.device_count() < ngpu:
        monkeypatch.setenv("MYQC_OVERSUBSCRIBE", "1")
    many = Q.eri_dense(s, ngpu=ngpu)
    assert np.abs(many - one).max() < 1e-13 and np.array_equal(many == 0.0, one == 0.0)
    packed = Q.eri_packed(s, ngpu=ngpu)
    assert np.abs(packed - Q.eri_packed(s, ngpu=1)).max() < 1e-13


def test_int2e_main_with_all_devices(tmp_path, oracle_inputs):
    """`int2e <ngpu>` / MYQC_NGPU: the drop-in program on every visible device (0 = all)."""
    s = product_system("CO2", tmp_path)
    assert Q.int2e_main(str(tmp_path), 0) == 0
    xx = Q.read_xx(str(tmp_path / "XX"), s.norb)
    mol, b, ft = oracle_system("CO2", oracle_inputs)
    ref, _ = O.int2e_dense(mol, b, ft)
    assert np.abs(xx - ref).max() < TOL


@pytest.mark.parametrize("name", ["h2o", "h2o_4", "c4h10", "h2o_8"])
def test_small_clusters_full_compare(name, tmp_path, oracle_inputs):
    s = product_system(name, tmp_path)
    got = Q.eri_packed(s)
    mol, b, ft = oracle_system(name, oracle_inputs)
    ref = O.int2e_packed(mol, b, ft)
    assert np.abs(got - ref).max() < TOL
    assert np.abs(got[ref == 0.0]).max(initial=0.0) < 1e-14


def test_h2o16_baseline_config_full_compare(tmp_path, oracle_inputs):
    """BASELINE.json configs[2]: (H2O)_16, 112 basis functions, every one of the 20 024 956 unique
    integrals against the oracle."""
    s = product_system("h2o_16", tmp_path)
    assert (s.norb, s.nunique) == (112, 20024956)
    got = Q.eri_packed(s)
    mol, b, ft = oracle_system("h2o_16", oracle_inputs)
    ref = O.int2e_packed(mol, b, ft)
    assert np.abs(got - ref).max() < TOL
    assert np.array_equal(got == 0.0, ref == 0.0) or np.abs(got[ref == 0.0]).max(initial=0.0) < 1e-14


def _stratified_rows(s, nrows, seed=20261017):
    """A fixed sample of packed rows P = (i,j) that covers every kind of row the kernels produce: rows of
    same-centre pairs (dense, every quartet class lands in them), rows of bonded / neighbouring pairs, rows of
    distant pairs (a handful of integrals), for s-s, s-p and p-p function pairs alike, plus the first and last row."""
    n = s.norb
    setl = int(s.setinfo[1])
    cen = np.zeros(n, dtype=int)
    lfn = np.zeros(n, dtype=int)
    for o in range(n):
        lfn[o] = int(s.basinfo[2 + 5 * o + 1])
        cen[o] = int(s.basinfo[2 + 5 * o + 4])
    xyz = np.array(s.xyz).reshape(3, s.nnuc).T
    rng = np.random.default_rng(seed)
    ii, jj = np.triu_indices(n)
    d = np.linalg.norm(xyz[cen[ii]] - xyz[cen[jj]], axis=1)
    kind = lfn[ii] + lfn[jj]                       # 0: s-s, 1: s-p, 2: p-p
    band = np.digitize(d, [1e-9, 3.0, 7.0, 12.0])  # same centre | bonded | near | mid | far (bohr)
    P = ii * n - ii * (ii - 1) // 2 + (jj - ii)
    rows = [0, s.npair - 1]
    per = max(1, nrows // 15)
    for k in range(3):
        for b in range(5):
            cand = P[(kind == k) & (band == b)]
            if len(cand):
                rows.extend(rng.choice(cand, size=min(per, len(cand)), replace=False).tolist())
    return np.unique(np.array(rows, dtype=np.int64))


def _rows_check(name, tmp_path, oracle_inputs, nrows):
    """Full-size configs: the whole packed array stays on the device side of the call; a fixed stratified
    sample of complete rows is compared with the (multithreaded) row oracle, plus size-independent properties."""
    s = product_system(name, tmp_path)
    got = Q.eri_packed(s)
    mol, b, ft = oracle_system(name, oracle_inputs)
    npair = s.npair
    rows = _stratified_rows(s, nrows)
    ref = O.int2e_rows(mol, b, ft, rows)

    def packed_index(P, Pp):
        lo, hi = np.minimum(P, Pp), np.maximum(P, Pp)
        return lo * npair - lo * (lo - 1) // 2 + (hi - lo)
    cols = np.arange(npair, dtype=np.int64)
    worst, nonempty, nnz = 0.0, 0, 0
    for r, P in enumerate(rows):
        g = got[packed_index(np.int64(P), cols)]
        worst = max(worst, float(np.abs(g - ref[r]).max()))
        assert np.array_equal(g == 0.0, ref[r] == 0.0) or np.abs(g[ref[r] == 0.0]).max(initial=0.0) < 1e-14
        k = int(np.count_nonzero(ref[r]))
        nnz += k
        nonempty += k > 0
    assert worst < TOL, worst
    # rows of distant pairs are (nearly) empty on purpose: they pin the exact zeros; most rows are not
    assert nonempty >= 0.5 * len(rows) and nnz > 50 * len(rows), (nonempty, len(rows), nnz)
    # Schwarz inequality |(P|P')| <= sqrt((P|P)(P'|P')) on the sampled rows (size independent).
    # Slack 1e-6: the reference's EIJ*EGH >= 1e-14 rule zeroes a diagonal (P'|P') once E_P' < 1e-7
    # while (P|P') with a compact P survives, so the inequality only holds up to ~E_P' itself.
    diag = got[packed_index(cols, cols)]
    assert diag.min() > -1e-12
    for P in rows[::4]:
        g = got[packed_index(np.int64(P), cols)]
        assert np.all(np.abs(g) <= np.sqrt(np.abs(diag[P]) * np.abs(diag)) + 1e-6)
    return s, got


def test_c20h42_rows(tmp_path, oracle_inputs):
    """BASELINE.json configs[3]: C20H42, 142 basis functions."""
    s, got = _rows_check("c20h42", tmp_path, oracle_inputs, 300)
    assert (s.norb, s.nunique) == (142, 51546781)


def test_h2o64_rows(tmp_path, oracle_inputs):
    """BASELINE.json configs[4]: (H2O)_64, 448 basis functions, 5.06e9 unique integrals (40.5 GB)."""
    s, got = _rows_check("h2o_64", tmp_path, oracle_inputs, 150)
    assert (s.norb, s.nunique) == (448, 5057816176)
    # translation invariance of the lattice: molecule (0,0,0) and molecule (1,0,0) have the same
    # intramolecular integrals (7 functions each, 16 molecules apart in orbital order)
    n, npair = s.norb, s.npair

    def idx(i, j, g, h):
        i, j = min(i, j), max(i, j)
        g, h = min(g, h), max(g, h)
        P, Pp = Q.pair_index(i, j, n), Q.pair_index(g, h, n)
        P, Pp = min(P, Pp), max(P, Pp)
        return P * npair - P * (P - 1) // 2 + (Pp - P)
    rng = np.random.default_rng(7)
    for _ in range(200):
        i, j, g, h = rng.integers(0, 7, 4)
        a = got[idx(i, j, g, h)]
        bshift = got[idx(i + 7 * 16, j + 7 * 16, g + 7 * 16, h + 7 * 16)]
        assert abs(a - bshift) < 1e-11


# ("CO2", 8), ("h2o", 7): more shards than shells, so some shards own no packed row at all (out_elems == 0)
@pytest.mark.parametrize("name,nsh", [("CO2", 2), ("h2o_4", 3), ("h2o_8", 4), ("h2o_16", 8), ("CO2", 8), ("h2o", 7)])
def test_shards_reproduce_the_unsharded_array(name, nsh, tmp_path):
    """Multi-GPU path on one device: every shard computes its slice independently.  A quartet that
    straddles two ownership lists is evaluated with bra and ket roles possibly exchanged with
    respect to the single-shard run (P-Q changes sign, sums run in another order), so the
    concatenation agrees to rounding (1e-13), not bit for bit; zeros stay exact zeros."""
    s = product_system(name, tmp_path)
    whole = Q.eri_packed(s)
    off = Q.shard_layout(s, nsh)
    parts = np.empty_like(whole)
    for k in range(nsh):
        Q.eri_packed_shard(s, parts[off[k]:off[k + 1]], device=0, shard=k, nshards=nsh)
    assert np.abs(parts - whole).max() < 1e-13
    assert np.array_equal(parts == 0.0, whole == 0.0)
    # and each run is deterministic
    again = np.empty_like(whole)
    for k in range(nsh):
        Q.eri_packed_shard(s, again[off[k]:off[k + 1]], device=0, shard=k, nshards=nsh)
    assert np.array_equal(parts, again)


def test_plan_api_device_buffers(tmp_path):
    import torch
    s = product_system("h2o_4", tmp_path)
    plan = Q.Plan(s, device=0)
    out = torch.full((plan.out_elems,), float("nan"), dtype=torch.float64, device="cuda:0")
    plan.execute(out.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    ref = Q.eri_packed(s)
    assert np.array_equal(out.cpu().numpy(), ref)  # every element written, deterministic
    ms = plan.execute_timed(out.data_ptr(), torch.cuda.current_stream().cuda_stream)
    assert len(ms) == len(plan.launches()) and all(m >= 0 for m in ms)
    st = plan.stats()
    assert sum(st["nquartets"]) == int(Q.canonical_stats(s)[0].sum())
    # dense expansion on device
    n = s.norb
    xx = torch.empty(n ** 4, dtype=torch.float64, device="cuda:0")
    Q._check(Q.lib().myqc_eri_expand_dense(out.data_ptr(), n, xx.data_ptr(), None))
    torch.cuda.synchronize()
    assert np.array_equal(xx.cpu().numpy().reshape((n,) * 4, order="F"), Q.eri_dense(s))
    plan.close()


@pytest.mark.parametrize("name", ["CO2", "h2o_8", "c4h10"])
def test_schwarz_skip_omits_less_than_tau_and_can_be_turned_off(name, tmp_path, monkeypatch):
    """The Schwarz skip (Q_u*Q_v < tau, Q from unscreened diagonals) leaves every omitted integral below tau;
    MYQC_SCHWARZ_TAU=0 evaluates exactly the reference's set.  The device counters count what the kernels evaluate:
    at tau = 0 at least the canonical primitive quartets of SURVEY.md 8d (a same-shell pair lists its primitive
    pairs in both orders, so they exceed the canonical count a little), fewer with the skip on."""
    import torch
    s = product_system(name, tmp_path)
    nq, _ = Q.canonical_stats(s)

    def run(tau):
        monkeypatch.setenv("MYQC_SCHWARZ_TAU", tau)
        plan = Q.Plan(s, device=0)
        out = torch.full((plan.out_elems,), float("nan"), dtype=torch.float64, device="cuda:0")
        plan.execute(out.data_ptr(), torch.cuda.current_stream().cuda_stream)
        ex, t = plan.executed_quartets()
        plan.close()
        return out.cpu().numpy(), ex, t
    exact, ex0, t0 = run("0")
    assert t0 == 0.0 and all(int(c) <= e <= 2 * int(c) + 64 for c, e in zip(nq, ex0))
    for tau in ("1e-12", "1e-11"):
        got, ex, t = run(tau)
        assert t == float(tau)
        assert np.abs(got - exact).max() < float(tau)
        assert all(a <= b for a, b in zip(ex, ex0))
        assert np.array_equal(got[exact == 0.0], exact[exact == 0.0])
    if name != "CO2":
        assert sum(ex) < sum(ex0)  # the last run (tau = 1e-11) did leave quartets out


def test_permuting_atoms_permutes_integrals(tmp_path, oracle_inputs):
    """Relabelling the nuclei only relabels the integrals (size-independent property)."""
    zm = example_zmat("h2o_2").splitlines()
    atoms = zm[1:7]
    perm_lines = [zm[0]] + atoms[3:] + atoms[:3] + zm[7:]
    s1 = Q.make_job(str(tmp_path / "a"), "\n".join(zm) + "\n", INPUTS)
    s2 = Q.make_job(str(tmp_path / "b"), "\n".join(perm_lines) + "\n", INPUTS)
    x1, x2 = Q.eri_dense(s1), Q.eri_dense(s2)
    p = np.concatenate([np.arange(7, 14), np.arange(0, 7)])
    assert np.abs(x1 - x2[np.ix_(p, p, p, p)]).max() < 1e-12


@pytest.mark.parametrize("nsh,shift,chunk", [(1, 0, None), (2, 1, None), (1, 1, 32), (2, 0, 64), (1, 3, 128), (2, 1, 256)])
def test_sparse_host_transfer_is_bit_identical(nsh, shift, chunk, tmp_path, monkeypatch):
    """Pinned destination: only the chunks that hold a nonzero cross PCIe (stored by the GPU into the host
    buffer), host threads write the zeros.  Same bytes as the plain cudaMemcpy path, for whole arrays and
    shards, at odd element offsets of the destination (the streaming-store zeroing has a scalar head and tail),
    for every chunk size, and with a poisoned destination (every element must be written by one side or the other)."""
    import torch
    s = product_system("h2o_16", tmp_path)
    off = Q.shard_layout(s, nsh)
    for sh in range(nsh):
        nloc = int(off[sh + 1] - off[sh])
        assert nloc >= 1 << 22  # large enough for the sparse path
        monkeypatch.setenv("MYQC_SPARSE_D2H", "0")
        plain = np.empty(nloc)
        Q.eri_packed_shard(s, plain, shard=sh, nshards=nsh)
        monkeypatch.delenv("MYQC_SPARSE_D2H")
        if chunk is not None:
            monkeypatch.setenv("MYQC_XFER_CHUNK", str(chunk))  # doubles per chunk (default 32 = one 256-byte warp store)
        pinned = torch.full((nloc + 4,), float("nan"), dtype=torch.float64).pin_memory()
        dst = pinned.numpy()[shift:shift + nloc]
        Q.eri_packed_shard(s, dst, shard=sh, nshards=nsh)
        assert np.array_equal(dst.view(np.int64), plain.view(np.int64))
        assert np.isnan(pinned.numpy()[shift + nloc])  # nothing written past the slice


@pytest.mark.parametrize("name,nsh", [("CO2", 1), ("c4h10", 1), ("h2o_8", 3), ("h2o_16", 1)])
def test_warp_cooperative_kernels_match_class_kernels(name, nsh, tmp_path, monkeypatch):
    """(SP SP|SP SP) and (S SP|SP SP) by the warp-cooperative kernel (a warp per contracted quartet, lanes over its
    primitive quartets, MYQC_PP_KERNEL / MYQC_SP_KERNEL = warp) against the class kernels (one lane per contracted
    quartet): the same screens and arithmetic, only the order in which the primitive quartets of one integral are added
    differs -- 1e-13 absolute, the same zero pattern; and the forced-warp array passes the oracle bar on CO2."""
    s = product_system(name, tmp_path)
    off = Q.shard_layout(s, nsh)

    def run():
        Q.release_cache()
        out = np.full(int(off[-1]), np.nan)
        for k in range(nsh):
            Q.eri_packed_shard(s, out[off[k]:off[k + 1]], device=0, shard=k, nshards=nsh)
        return out
    monkeypatch.setenv("MYQC_PP_KERNEL", "slices")
    monkeypatch.setenv("MYQC_SP_KERNEL", "class")
    ref = run()
    monkeypatch.setenv("MYQC_PP_KERNEL", "warp")
    monkeypatch.setenv("MYQC_SP_KERNEL", "warp")
    got = run()
    Q.release_cache()
    assert not np.isnan(got).any()
    assert np.abs(got - ref).max() < 1e-13
    assert np.array_equal(got == 0.0, ref == 0.0)


@pytest.mark.parametrize("engine", ["kernel", "copy"])
@pytest.mark.parametrize("name,nsh", [("CO2", 1), ("h2o_8", 3), ("h2o_16", 1)])
def test_fill_engines_agree(name, nsh, engine, tmp_path, monkeypatch):
    """The zero fill of the slice by cudaMemsetAsync (default), by the repo's fill kernel (MYQC_FILL_ENGINE=kernel) and
    by device-to-device copies from a zero buffer on several streams (=copy) give the same bytes."""
    s = product_system(name, tmp_path)
    off = Q.shard_layout(s, nsh)

    def run():
        Q.release_cache()
        out = np.full(int(off[-1]), np.nan)
        for k in range(nsh):
            Q.eri_packed_shard(s, out[off[k]:off[k + 1]], device=0, shard=k, nshards=nsh)
        return out
    ref = run()
    monkeypatch.setenv("MYQC_FILL_ENGINE", engine)
    monkeypatch.setenv("MYQC_FILL_STREAMS", "3")
    monkeypatch.setenv("MYQC_ZERO_MB", "1")
    got = run()
    Q.release_cache()
    assert np.array_equal(got.view(np.int64), ref.view(np.int64))


@pytest.mark.parametrize("name,nsh", [("CO2", 1), ("h2o_8", 1), ("c4h10", 1), ("h2o_8", 3), ("CO2", 8), ("h2o_16", 1)])
def test_compose_mode_equals_scatter_mode(name, nsh, tmp_path, monkeypatch):
    """MYQC_OUTPUT_MODE=compose (class kernels stage dense quartet blocks, compose_kernel writes every element of the
    slice once) produces the array of the default scatter mode bit for bit: the same lanes evaluate the same quartets
    in the same order, only the way the integrals reach the packed array differs."""
    s = product_system(name, tmp_path)
    off = Q.shard_layout(s, nsh)

    def run():
        Q.release_cache()
        out = np.full(int(off[-1]), np.nan)
        for k in range(nsh):
            Q.eri_packed_shard(s, out[off[k]:off[k + 1]], device=0, shard=k, nshards=nsh)
        return out
    ref = run()
    monkeypatch.setenv("MYQC_OUTPUT_MODE", "compose")
    got = run()
    Q.release_cache()
    assert not np.isnan(got).any()
    assert np.array_equal(got, ref)
