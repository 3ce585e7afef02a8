"""Result files in the reference's HDF5 layout (/root/reference/src/IO.jl:64-143,
ObservablesGeneric.jl:49-72): the pure-NumPy writer, its independent reader, and the getmeans / getenergies post-processing."""
import struct

import numpy as np
import pytest


def test_round_trip_of_every_supported_kind(scmc, tmp_path):
    h5 = scmc.hdf5
    rng = np.random.default_rng(5)
    tree = {"current_sweep": 20000, "σ": 0.0275, "energy": rng.normal(size=37), "βs": np.zeros(0),
            "state": rng.normal(size=(9, 4)) + 1j * rng.normal(size=(9, 4)), "rs": rng.normal(size=(18, 2)),
            "bytes": np.arange(7, dtype=np.uint8), "i32": np.arange(6, dtype=np.int32).reshape(2, 3), "f32": np.float32(1.5),
            "means": {"energy": {"mean": -3.54, "error": 9e-5}, "στ": {"mean": 0.1, "error": 0.01},
                      "magnetizationVec": {"mean": rng.normal(size=16), "error": rng.random(16)}},
            "many": {f"k{i:02d}": float(i) for i in range(23)}}         # more entries than one default symbol-table node holds
    fn = tmp_path / "a.h5"
    h5.write(fn, tree)
    back = h5.read(fn)

    def same(a, b):
        if isinstance(a, dict):
            assert sorted(a) == sorted(b)
            for k in a:
                same(a[k], b[k])
        else:
            a = np.asarray(a)
            assert np.asarray(b).shape == a.shape and np.asarray(b).dtype.kind == a.dtype.kind
            assert np.array_equal(np.asarray(b), a)
    same(tree, back)
    assert back["f32"].dtype == np.float32 and back["i32"].dtype == np.int32 and back["current_sweep"].dtype == np.int64


def test_file_structure_by_the_specification(scmc, tmp_path):
    """Known answers at the byte level: format signature and superblock v0 fields, the root symbol-table entry, the structure
    signatures found at the addresses it names, node sizes, and the messages of a float64 dataset header."""
    h5 = scmc.hdf5
    fn = tmp_path / "b.h5"
    h5.write(fn, {"x": np.array([1.0, 2.0, 3.0])})
    d = open(fn, "rb").read()
    assert d[:8] == bytes([0x89, 0x48, 0x44, 0x46, 0x0D, 0x0A, 0x1A, 0x0A])
    assert d[8:16] == bytes([0, 0, 0, 0, 0, 8, 8, 0])                   # versions 0, 8-byte offsets and lengths
    assert struct.unpack_from("<HHI", d, 16) == (4, 16, 0)              # leaf K, internal K, consistency flags
    base, free, eof, drv = struct.unpack_from("<4Q", d, 24)
    assert (base, free, eof, drv) == (0, 2 ** 64 - 1, len(d), 2 ** 64 - 1)
    name0, oh, cache, _, bt, hp = struct.unpack_from("<QQIIQQ", d, 56)
    assert name0 == 0 and cache == 1 and oh % 8 == bt % 8 == hp % 8 == 0
    assert d[bt:bt + 4] == b"TREE" and d[hp:hp + 4] == b"HEAP"
    assert struct.unpack_from("<BBH", d, bt + 4) == (0, 0, 1)           # group node, leaf level, one child
    assert hp - bt == 24 + 33 * 8 + 32 * 8                              # a full node of 2K+1 keys and 2K children is on disk
    seg_size, free_head, seg = struct.unpack_from("<QQQ", d, hp + 8)
    assert seg == hp + 32 and free_head < seg_size and d[seg] == 0 and d[seg + 8:seg + 10] == b"x\0"
    assert struct.unpack_from("<QQ", d, seg + free_head) == (1, seg_size - free_head)   # the free block ends the list and the segment
    key0, snod, key1 = struct.unpack_from("<QQQ", d, bt + 24)
    assert (key0, key1) == (0, 8) and d[snod:snod + 4] == b"SNOD" and struct.unpack_from("<BBH", d, snod + 4) == (1, 0, 1)
    off, doh, dcache = struct.unpack_from("<QQI", d, snod + 8)
    assert off == 8 and dcache == 0
    ver, _, nmsg, refs, size = struct.unpack_from("<BBHII", d, doh)
    assert (ver, nmsg, refs) == (1, 4, 1)
    p, seen = doh + 16, {}
    for _ in range(nmsg):
        t, n, _ = struct.unpack_from("<HHB", d, p)
        assert n % 8 == 0
        seen[t] = d[p + 8:p + 8 + n]
        p += 8 + n
    assert p == doh + 16 + size
    assert seen[1][:2] == bytes([1, 1]) and struct.unpack_from("<Q", seen[1], 8) == (3,)
    # IEEE double, little endian: class 1 version 1, mantissa normalisation "implied", sign bit 63, 8 bytes;
    # precision 64, exponent at 52 (11 bits, bias 1023), mantissa at 0 (52 bits)
    assert seen[3][:20] == bytes([0x11, 0x20, 63, 0, 8, 0, 0, 0, 0, 0, 64, 0, 52, 11, 0, 52, 0xFF, 0x03, 0, 0])
    assert seen[8][:2] == bytes([3, 1])
    addr, nbytes = struct.unpack_from("<QQ", seen[8], 2)
    assert nbytes == 24 and np.array_equal(np.frombuffer(d, "<f8", 3, addr), [1.0, 2.0, 3.0])
    with pytest.raises(ValueError):
        open(tmp_path / "c.h5", "wb").write(d[:100])
        h5.read(tmp_path / "c.h5")


def test_complex_is_the_compound_of_hdf5_jl(scmc, tmp_path):
    h5 = scmc.hdf5
    fn = tmp_path / "z.h5"
    z = np.array([[1 + 2j, 3 - 4j]])
    h5.write(fn, {"state": z})
    d = open(fn, "rb").read()
    i = d.index(bytes([0x16, 2, 0, 0, 16, 0, 0, 0]))                    # compound, version 1, two members, 16 bytes
    assert d[i + 8:i + 16] == b"r" + b"\0" * 7 and struct.unpack_from("<I", d, i + 16) == (0,)
    j = i + 8 + 8 + 32 + 20
    assert d[j:j + 8] == b"i" + b"\0" * 7 and struct.unpack_from("<I", d, j + 8) == (8,)
    assert np.array_equal(h5.read(fn)["state"], z)


def test_update_keeps_existing_entries(scmc, tmp_path):
    h5 = scmc.hdf5
    fn = tmp_path / "u.h5"
    h5.write(fn, {"energy": np.arange(3.0), "means/energy/mean": 1.0})
    h5.update(fn, {"means/energy/error": 0.5, "rs": np.ones((2, 2))})
    back = h5.read(fn)
    assert sorted(back) == ["energy", "means", "rs"] and back["means"]["energy"] == {"mean": 1.0, "error": 0.5}


class _Cfg:
    def __init__(self, N=6, d=4, R=3):
        rng = np.random.default_rng(1)
        self.positions = rng.normal(size=(N, 2)); self.basis = [np.zeros(2)]
        self.seed, self.sweep_counter, self.replica_offset, self.nSites = 7, 40, 10, N
        self._psi = rng.normal(size=(R, N, d)) + 1j * rng.normal(size=(R, N, d))

    def get_state(self):
        return self._psi


class _Obs:
    def __init__(self, shift):
        self.shift = shift

    def means(self, cfg, beta, replica=0):
        x = self.shift + replica
        return {"energy": (-3.0 + x, 1e-3), "heat": (beta * x, 0.1), "magnetizationVec": (np.full(16, x), np.full(16, 0.01)),
                "correlations": (np.arange(16 * 6, dtype=float).reshape(16, 6) + x, np.ones((16, 6)))}


def test_result_files_and_post_processing(scmc, tmp_path):
    """saveMeans! + checkpoint! per temperature, then getmeans / getenergies over the temperature grid (IO.jl:64-143)."""
    io = scmc.io
    cfg = _Cfg()
    Ts = [0.5, 1.0, 2.0]
    name = lambda T: str(tmp_path / f"run_T{T}.h5")
    for r, T in enumerate(Ts):
        io.checkpointH5_(name(T), cfg, 40, [1 / T] * 4, np.arange(4.0) + r, [0.1, 0.2, 0.3], replica=r)
        io.saveMeans_(name(T), _Obs(0.0), cfg, 1 / T, replica=r)
    raw = scmc.hdf5.read(name(1.0))
    assert raw["state"].shape == (6, 4) and raw["rs"].shape == (6, 2)                 # HDF5 dims = reversed Julia dims (d x N, dim x |rs|)
    assert raw["means"]["correlations"]["mean"].shape == (6, 16)
    assert raw["σ"] == 0.2 and raw["current_sweep"] == 40 and raw["replica_id"] == 11
    one = io.readCheckpointH5(name(1.0))
    assert np.array_equal(one["state"], cfg.get_state()[1].T) and one["means"]["correlations"]["mean"].shape == (16, 6)
    out = io.getmeansH5(name, str(tmp_path / "means.h5"), Ts + [3.0])                  # the last file does not exist
    assert out["energy/mean"].shape == (4,) and np.allclose(out["energy/mean"], [-3.0, -2.0, -1.0, 0.0])
    assert out["correlations/mean"].shape == (16, 6, 4) and out["correlations/mean"][2, 3, 1] == 2 * 6 + 3 + 1
    m = io.readCheckpointH5(str(tmp_path / "means.h5"))
    assert np.array_equal(m["Ts"], Ts + [3.0]) and np.array_equal(m["magnetizationVec"]["mean"], out["magnetizationVec/mean"])
    assert scmc.hdf5.read(str(tmp_path / "means.h5"))["correlations"]["mean"].shape == (4, 6, 16)
    en = io.getenergiesH5(name, str(tmp_path / "energies.h5"), Ts + [3.0])
    back = scmc.hdf5.read(str(tmp_path / "energies.h5"))
    assert sorted(back) == ["0.5", "1.0", "2.0", "3.0"] and np.array_equal(back["2.0"], np.arange(4.0) + 2) and np.array_equal(back["3.0"], [0.0])
    assert set(en) == set(back)


def test_result_file_names_of_a_packed_launch(scmc):
    """One file per chain: `filename(T)` as in notebook cell 22, or a path that gets the temperature (and the replica index when a
    temperature occurs more than once); names that are not .h5 paths are labels for the log and write nothing."""
    f = scmc.launcher.resultFilename
    assert f("out/run.h5", 0.5, None, 1) == "out/run.h5"
    assert f("out/run.h5", 0.5, None, 3) == "out/run_T0.5.h5" and f("out/run.h5", 0.02, 4, 8) == "out/run_T0.02_r4.h5"
    assert f(lambda T: f"x{T}.h5", 2.0, None, 5) == "x2.0.h5"
    w = scmc.launcher._writes_files
    assert w("a.h5") and w("b.HDF5") and w(lambda T: "c") and not w("label") and not w(None)


def _fnv(raw):
    s = 0
    for b in raw:
        s = (s * 1099511628211 + b) & 0xFFFFFFFFFFFFFFFF
    return f"{s:016x}"


def test_files_pass_an_independent_c_reader(tmp_path):
    """No libhdf5 / h5py / HDF5.jl exists in this image, so the written files are checked by a SECOND reader, written in C from the HDF5 File
    Format Specification alone (tests/h5_check.c: superblock v0, v1 object headers, symbol-table groups through B-tree / SNOD / local heap,
    dataspace, datatype and contiguous-layout messages, address and size consistency).  Its listing must name every object with the type,
    shape, leading values and a checksum of the raw bytes that were written; a damaged file must fail its checks."""
    import os
    import subprocess
    import sys
    ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, ROOT)
    import scmc_b200 as scmc
    exe = str(tmp_path / "h5_check")
    subprocess.check_call(["gcc", "-O1", "-Wall", "-Werror", "-o", exe, os.path.join(ROOT, "tests", "h5_check.c")])
    state = (np.arange(24).reshape(6, 4) + 1j * np.arange(24).reshape(6, 4)[::-1]).astype(np.complex128)
    tree = {"current_sweep": np.int64(200), "energy": np.linspace(-3.5, -3.4, 7), "σ": np.float64(0.0271), "βs": np.full(7, 50.0), "state": state,
            "means/energy/mean": np.float64(-3.54), "means/energy/error": np.float64(9e-5), "means/magnetizationVec/mean": np.arange(16.0),
            "means/correlations/mean": np.arange(16.0 * 36).reshape(36, 16), "rs": np.arange(72.0).reshape(36, 2), "seed": np.uint64(2 ** 63 + 5),
            "small": np.arange(5, dtype=np.int32) - 2, "f32": np.arange(3, dtype=np.float32) / 3, "empty": np.zeros(0)}
    tree.update({f"means/obs{k:02d}/mean": np.float64(k) for k in range(20)})          # a group larger than the default leaf K
    fn = str(tmp_path / "t.h5")
    scmc.hdf5.write(fn, tree)
    out = subprocess.run([exe, fn], capture_output=True, text=True, timeout=60)
    assert out.returncode == 0 and out.stdout.rstrip().endswith("h5_check ok"), out.stdout + out.stderr
    lines = {l.split(" ", 1)[0]: l.split(" ", 1)[1] for l in out.stdout.splitlines()[1:-1]}
    names = {"int64": "int64", "int32": "int32", "uint64": "uint64", "float64": "float64", "float32": "float32", "complex128": "complex128"}
    n_sets = 0
    for key, v in tree.items():
        a = np.asarray(v)
        got = lines["/" + key]
        shape = "[" + ",".join(str(n) for n in a.shape) + "]"
        assert got.startswith(f"{names[a.dtype.name]} {shape}"), (key, got)
        assert got.endswith("fnv=" + _fnv(a.tobytes())), (key, got)
        n_sets += 1
    groups = [k for k, v in lines.items() if v == "group"]
    assert sorted(groups) == sorted(["/", "/means/", "/means/energy/", "/means/magnetizationVec/", "/means/correlations/"] + [f"/means/obs{k:02d}/" for k in range(20)])
    assert len(lines) == n_sets + len(groups)
    assert lines["/state"].split()[2:4] == ["(0,20)", "(1,21)"] and lines["/small"].split()[2:7] == ["-2", "-1", "0", "1", "2"]
    assert lines["/seed"].split()[2] == str(2 ** 63 + 5) and float(lines["/σ"].split()[2]) == 0.0271
    # the reference's result file as launch_ writes it goes through the same checks on the GPU box (tests/test_gpu_configs.py); here: damage
    raw = bytearray(open(fn, "rb").read())
    raw[40] ^= 0x10                                                                     # end-of-file address in the superblock
    bad = str(tmp_path / "bad.h5")
    open(bad, "wb").write(raw)
    assert subprocess.run([exe, bad], capture_output=True, text=True, timeout=60).returncode == 1
    raw = bytearray(open(fn, "rb").read())
    k = raw.find(b"SNOD")
    raw[k + 8 + 40:k + 8 + 48], raw[k + 8:k + 16] = raw[k + 8:k + 16], raw[k + 8 + 40:k + 8 + 48]      # swap two link-name offsets: entries no longer sorted
    open(bad, "wb").write(raw)
    assert subprocess.run([exe, bad], capture_output=True, text=True, timeout=60).returncode == 1
