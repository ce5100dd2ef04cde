"""Host-side logic of the product, no GPU needed: the C-ABI library loads and exports what
include/myqc_eri.h declares, the file layer restates getenv/buildBasis/Ftab/XX exactly, the
parse stage matches the oracle's restatement, sharding tiles the packed array."""
import os
import re
import struct

import numpy as np
import pytest

import myqc_b200 as Q
from myqc_b200 import molecules, parse
from conftest import EXAMPLES, INPUTS, ROOT, example_zmat, oracle_system, product_system
from oracle import oracle as O


def test_library_exports_every_declared_symbol():
    import glob
    hdr = "".join(open(h).read() for h in sorted(glob.glob(os.path.join(ROOT, "include", "*.h"))))
    declared = set(re.findall(r"\b(myqc_[a-z0-9_]+)\s*\(", hdr))
    L = Q.lib()
    missing = [n for n in declared if not hasattr(L, n)]
    assert not missing, missing
    assert declared == set(Q.EXPORTS)


def test_no_cpu_fallback(tmp_path):
    """Without a device the compute entry points fail loudly instead of falling back."""
    if Q.device_count() > 0:
        pytest.skip("a GPU is present")
    s = product_system("H2", tmp_path)
    with pytest.raises(Q.MyQCError) as e:
        Q.eri_packed(s)
    assert e.value.code == Q.ERR_NO_DEVICE
    with pytest.raises(Q.MyQCError):
        Q.Plan(s)
    with pytest.raises(Q.MyQCError) as e:
        Q.fock_rhf(np.zeros(1), 1, np.zeros((1, 1)))
    assert e.value.code == Q.ERR_NO_DEVICE
    with pytest.raises(Q.MyQCError) as e:
        Q.ao2mo_transform(np.zeros(1), 1, *[np.ones((1, 1))] * 4)
    assert e.value.code == Q.ERR_NO_DEVICE


@pytest.mark.parametrize("name", EXAMPLES + ["h2o_4", "c4h10"])
def test_input_layer_matches_oracle(name, tmp_path, oracle_inputs):
    s = product_system(name, tmp_path)
    mol, b, ft = oracle_system(name, oracle_inputs)
    assert np.array_equal(s.atoms, mol.atoms)
    assert np.array_equal(s.xyz, mol.xyz_fortran())          # parse: COM shift + A2B, bit for bit
    assert np.array_equal(s.set, b.set) and np.array_equal(s.bas, b.bas)
    assert np.array_equal(s.setinfo, b.setinfo) and np.array_equal(s.basinfo, b.basinfo)
    assert np.array_equal(s.ftab, ft)
    # basinfo text file: downstream stages read tokens 0-1 = OpS, norb (RHFI2G.f90:48-51)
    Q.build_basis(os.path.join(INPUTS, "mybasis"), s.atoms, 0, out_dir=str(tmp_path))
    toks = open(tmp_path / "basinfo").read().split()
    assert int(toks[0]) == 4 and int(toks[1]) == s.norb
    stoks = open(tmp_path / "setinfo").read().split()
    assert int(stoks[0]) == s.nset and int(stoks[1]) == 7


def test_sizes_of_baseline_configs(tmp_path):
    for name, norb, nset in [("CO2", 15, 18), ("HF", 6, 9), ("CO", 10, 12), ("NO", 10, 12),
                             ("h2o_16", 112, 192), ("c20h42", 142, 246), ("h2o_64", 448, 768)]:
        atoms, xyz, opts = parse.parse_zmat(example_zmat(name))
        set_, setinfo, bas, basinfo, _, _ = Q.build_basis(os.path.join(INPUTS, "mybasis"), atoms)
        assert (basinfo[1], setinfo[0]) == (norb, nset)


def test_canonical_work_counts_match_survey(tmp_path):
    """SURVEY.md 8d work totals (exact canonical primitive-quartet counts and model flops)."""
    want = {"CO2": (11901, 6.462e6), "HF": (1035, 2.50e5), "CO": (2823, 1.455e6), "NO": (2714, 1.440e6),
            "h2o_16": (24634506, 3.520e9), "c20h42": (39864177, 6.748e9), "h2o_64": (1071129508, 1.397e11)}
    for name, (nq, fl) in want.items():
        s = product_system(name, tmp_path / name)
        got, flops = Q.canonical_stats(s)
        assert int(got.sum()) == nq
        assert abs(flops - fl) / fl < 2e-3


def test_ftab_reader_rejects_garbage(tmp_path):
    bad = tmp_path / "Ftab"
    bad.write_bytes(b"\x00" * 100)
    with pytest.raises(Q.MyQCError) as e:
        Q.read_ftab(str(bad))
    assert e.value.code == Q.ERR_IO
    with pytest.raises(Q.MyQCError):
        Q.read_ftab(str(tmp_path / "missing"))


def test_xx_record_framing(tmp_path):
    """One Fortran unformatted sequential record; gfortran subrecords above the limit: leading
    marker negative when continued, trailing marker negative when it has a predecessor."""
    n = 5
    xx = np.arange(n ** 4, dtype=np.float64).reshape((n,) * 4, order="F")
    p = str(tmp_path / "XX")
    Q.write_xx(p, xx, n)
    raw = open(p, "rb").read()
    assert len(raw) == 8 * n ** 4 + 8
    assert struct.unpack("<i", raw[:4])[0] == 8 * n ** 4 == struct.unpack("<i", raw[-4:])[0]
    assert np.array_equal(np.frombuffer(raw[4:-4], dtype="<f8"), xx.reshape(-1, order="F"))
    assert np.array_equal(Q.read_xx(p, n), xx)
    # force three subrecords of at most 2000 bytes (5000-byte payload)
    Q.write_xx(p, xx, n, max_subrecord=2000)
    raw = open(p, "rb").read()
    assert len(raw) == 8 * n ** 4 + 3 * 8
    marks, pos = [], 0
    while pos < len(raw):
        lead = struct.unpack("<i", raw[pos:pos + 4])[0]
        size = abs(lead)
        trail = struct.unpack("<i", raw[pos + 4 + size:pos + 8 + size])[0]
        marks.append((lead, trail))
        pos += 8 + size
    assert marks == [(-2000, 2000), (-2000, -2000), (1000, -1000)]
    assert np.array_equal(Q.read_xx(p, n), xx)


def test_int2e_main_error_paths(tmp_path):
    """Missing inputs -> `touch error`, non-zero status, no XX (int2e.f90:174-178, env.f90:37-53)."""
    rc = Q.int2e_main(str(tmp_path), 1)
    assert rc == Q.ERR_IO and (tmp_path / "error").exists() and not (tmp_path / "XX").exists()
    # existing XX -> skipped untouched (int2e.f90:58-63)
    d = tmp_path / "job"
    s = product_system("H2", d)
    (d / "XX").write_bytes(b"keep")
    assert Q.int2e_main(str(d), 1) == 0
    assert (d / "XX").read_bytes() == b"keep" and not (d / "error").exists()


def test_unsupported_inputs_are_rejected(tmp_path):
    s = product_system("HF", tmp_path)
    bad = np.array(s.setinfo, copy=True)
    bad[0] = s.nset + 1  # header does not match
    with pytest.raises(Q.MyQCError) as e:
        Q.shard_layout(Q.System(s.nnuc, s.atoms, s.xyz, s.set, bad, s.bas, s.basinfo, s.ftab), 2)
    assert e.value.code == Q.ERR_BAD_ARG
    bi = np.array(s.basinfo, copy=True)
    bi[1 + 5 * 1 + 2] = 2  # a d function
    with pytest.raises(Q.MyQCError) as e:
        Q.shard_layout(Q.System(s.nnuc, s.atoms, s.xyz, s.set, s.setinfo, s.bas, bi, s.ftab), 2)
    assert e.value.code == Q.ERR_UNSUPPORTED


@pytest.mark.parametrize("name,nsh", [("CO2", 2), ("h2o_4", 3), ("h2o_16", 8), ("c20h42", 4), ("h2o_64", 8)])
def test_shard_layout_tiles_the_packed_array(name, nsh, tmp_path):
    s = product_system(name, tmp_path)
    off = Q.shard_layout(s, nsh)
    assert off[0] == 0 and off[-1] == s.nunique and np.all(np.diff(off) >= 0)
    # every cut is the start of a packed row whose leading orbital starts a shell
    n, npair = s.norb, s.npair
    starts = {0, s.nunique}
    for i in range(n):
        P = i * n - i * (i - 1) // 2
        starts.add(P * npair - P * (P - 1) // 2)
    assert all(int(o) in starts for o in off)
    if name in ("h2o_16", "h2o_64"):
        assert np.count_nonzero(np.diff(off)) == nsh  # no empty shard on the benchmark workloads


@pytest.mark.parametrize("name,nsh", [("h2o_16", 4), ("c20h42", 4), ("h2o_64", 8)])
def test_shard_cost_model_is_consistent_and_balanced(name, nsh, tmp_path):
    """Host only (myqc_eri_shard_model): the seconds the cut model expects of every shard.  The fill part is the bytes
    of the shard's slice over the fill rate; the class part of all shards adds up to the canonical primitive-quartet
    counts times the per-class constants (the attribution to blocks only distributes it); and the cuts balance the
    totals as well as cuts at shell starts allow."""
    s = product_system(name, tmp_path)
    cs, fs = Q.shard_model(s, nsh)
    off = Q.shard_layout(s, nsh)
    assert cs.shape == (nsh,) and fs.shape == (nsh,) and np.all(cs >= 0) and np.all(fs >= 0)
    assert np.allclose(fs, 8.0 * np.diff(off) / 7.3e12, rtol=1e-12, atol=0)
    nq, _ = Q.canonical_stats(s)
    keff = np.array([0.46, 0.47, 0.32, 0.36, 0.30, 0.27])
    want = float(np.sum(nq * np.array(Q.CLASS_W) / (keff * 34.2e12))) * (1.0 + 0.07 * np.log2(nsh))
    assert abs(cs.sum() - want) < 1e-9 * want
    one_c, one_f = Q.shard_model(s, 1)
    assert abs(one_f[0] - 8.0 * s.nunique / 7.3e12) < 1e-15
    tot = cs + fs
    if name == "h2o_64":
        assert tot.max() / tot.mean() < 1.05  # the benchmark workload: within 5 % of the mean in the model


def test_streaming_host_zero_writes_exactly_its_range():
    """The zeroing routine of the sparse device -> host route (non-temporal stores for whole cache lines, scalar head
    and tail): every offset relative to a cache line, lengths around its thresholds, nothing outside the range."""
    buf = np.empty(4096 + 64)
    base = (-buf.ctypes.data // 8) % 8  # element index of a 64-byte boundary
    for off in range(0, 9):
        for n in [0, 1, 7, 8, 9, 31, 32, 33, 63, 64, 65, 255, 256, 257, 1000, 4000]:
            buf[:] = np.nan
            Q.host_zero(buf[base + off:base + off + n])
            assert np.all(buf[base + off:base + off + n] == 0.0) and not np.signbit(buf[base + off:base + off + n]).any()
            assert np.isnan(buf[:base + off]).all() and np.isnan(buf[base + off + n:]).all()


def test_synthetic_geometries(tmp_path):
    """deterministic lattices, atom order O,H,H, 3.0 A spacing."""
    atoms, xyz, _ = parse.parse_zmat(molecules.zmat("h2o_16"))
    assert len(atoms) == 48 and list(atoms[:3]) == [8, 1, 1]
    d = np.linalg.norm(xyz[3] - xyz[0]) / parse.A2B
    assert abs(d - 3.0) < 1e-12
    atoms, xyz, _ = parse.parse_zmat(molecules.zmat("c20h42"))
    assert (atoms == 6).sum() == 20 and (atoms == 1).sum() == 42
    cc = np.linalg.norm(xyz[1] - xyz[0]) / parse.A2B
    assert abs(cc - 1.54) < 1e-6


def test_pack_dense_helper_matches_the_numpy_packing():
    """myqc_pack_dense (what `ao2mo` applies to the XX record it reads) == the numpy triangle gather."""
    rng = np.random.default_rng(4)
    n = 5
    a = rng.standard_normal((n, n, n, n))
    xx = a + a.transpose(1, 0, 2, 3)
    xx = xx + xx.transpose(0, 1, 3, 2)
    xx = xx + xx.transpose(2, 3, 0, 1)
    assert np.array_equal(Q.pack_dense_c(xx), Q.pack_dense(xx))
    assert np.array_equal(Q.pack_dense_c(xx), O.packed_from_dense(xx))


def test_ao2mo_program_without_a_device_touches_error(tmp_path):
    """PROGRAM ao2mo has no CPU fallback either: with the inputs in place but no GPU it reports and
    leaves the `error` sentinel the driver tests (myQC.f90:74-78)."""
    if Q.device_count() > 0:
        pytest.skip("a GPU is present")
    Q.make_job(str(tmp_path), example_zmat("Be"), INPUTS)  # CALC= MP2, REF= RHF
    open(tmp_path / "basinfo", "w").write(" 4 5\n")
    assert Q.ao2mo_main(str(tmp_path)) == Q.ERR_NO_DEVICE and (tmp_path / "error").exists()


# ---- N3: the `parse` stage in C++ (include/myqc_parse.h) ----------------------------------------
@pytest.mark.parametrize("name", EXAMPLES + ["O_triplet", "h2o_16", "c20h42", "h2o_64"])
def test_cpp_parser_matches_python_mirror_and_oracle(name):
    """myqc_parse_zmat == myqc_b200.parse == the oracle's restatement of parser.f90: atoms, the
    centre-of-mass / unit-converted geometry bit for bit, all 17 options, electron counts."""
    zm = example_zmat(name)
    a, x, o, na, nb, problems = Q.parse_zmat(zm)
    a2, x2, o2 = parse.parse_zmat(zm)
    m = O.parse_zmat(zm)
    assert np.array_equal(a, a2) and np.array_equal(x, x2) and np.array_equal(o, o2)
    assert np.array_equal(x, m.xyz) and np.array_equal(a, m.atoms)
    assert (na, nb) == parse.electron_counts(a2, o2) == tuple(O.electrons(m)) and problems == 0


def test_parse_program_files_and_error_sentinel(tmp_path):
    """PROGRAM parser at the process boundary: nucpos / envdat / fmem token for token as the Python mirror
    writes them; `error` in the cases the reference touches it (parser.f90:81-84, 640-644, 513-524, 94-99)."""
    def run(zm, sub):
        d = tmp_path / sub
        d.mkdir()
        (d / "ZMAT").write_text(zm)
        return d, Q.parse_main(str(d))

    zm = example_zmat("OH")
    d, rc = run(zm, "ok")
    assert rc == 0 and not (d / "error").exists()
    ref = tmp_path / "ref"
    parse.write_job_files(str(ref), *parse.parse_zmat(zm))
    for f in ("nucpos", "envdat", "fmem"):
        assert (d / f).read_text().split() == (ref / f).read_text().split(), f
    atoms, xyz, na, nb, fmem, opts = Q.read_env(str(d))
    assert (na, nb) == (5, 4) and fmem == 1000.0 and opts[13] == 1 and opts[15] == 4  # EXCITE= CIS, E_NUM= 4
    # the record right after END is skipped by read_options' bare READ(1,*) -- also when it is an option line
    d, rc = run(zm.replace("END\n\n", "END\n"), "skipped")
    assert rc == 0 and Q.read_env(str(d))[5][1] == 0 and parse.parse_zmat(zm.replace("END\n\n", "END\n"))[2][1] == 0
    d, rc = run(zm.replace("CARTESIAN", "INTERNAL"), "internal")
    assert rc != 0 and (d / "error").exists() and not (d / "nucpos").exists()
    d, rc = run(zm.replace("END", ""), "noend")
    assert rc != 0 and (d / "error").exists()
    d, rc = run(zm.replace("REF= UHF", "REF= RHF"), "rhf_open_shell")  # files are written, error is touched
    assert (d / "error").exists() and (d / "envdat").exists()
    d, rc = run(example_zmat("H2").replace("0.50", "0.10"), "close")
    assert (d / "error").exists()
    d, rc = run(example_zmat("CO2").replace("CALC= MP2", "CALC= MP2\nEXCITE= CIS"), "bad_options")
    assert rc != 0 and (d / "error").exists()
    d, rc = run(example_zmat("HF") + "FOO= 1\n", "unknown_key")  # reported, ignored
    assert rc == 0 and not (d / "error").exists()
    assert Q.parse_main(str(tmp_path / "nowhere")) == Q.ERR_IO


def test_bench_reference_arm_prints_one_json_line_with_the_contract_keys():
    """`bench.py --impl reference` (the CPU arm the driver runs next to ours) on the smallest workload:
    exactly one line on stdout, valid JSON, the keys the contract names."""
    import json
    import subprocess
    import sys
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "CO2",
                          "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, out.stdout[:2000]
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["metric"] == "unique_eris_per_s" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["config"]["workload"] == "CO2"
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0


def test_bench_reference_arm_under_torchrun_only_rank0_prints():
    """The driver launches the reference arm like ours (torchrun, N ranks): rank 0 prints the line, the others
    exit 0 without output."""
    import json
    import subprocess
    import sys
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "bench.py"),
                          "--impl", "reference", "--gpus", "2", "--workload", "CO2", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip().startswith("{")]
    assert len(lines) == 1, out.stdout[:2000]
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 2
