"""CPU tests of the drop-in boundary: the shared library loads without a GPU, exports every symbol the headers
declare, serves the host-side (init-time) functions, and FAILS LOUDLY -- instead of falling back to the CPU -- when
asked to compute without a device."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

import srsran_b200 as b
from util import all_K, lanes8, lanes16

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from srsran_b200 import build
    build.build()
    return b.lib()


def declared_symbols():
    names = []
    for h in ("fec.h", "batch.h"):
        txt = open(os.path.join(ROOT, "include", "srslte_b200", h)).read()
        txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
        for m in re.finditer(r"SRSLTE(?:_B200)?_API\s+[\w\s\*]+?\b(srslte_\w+)\s*\(", txt):
            names.append(m.group(1))
    return sorted(set(names))


def test_every_declared_symbol_is_exported(lib):
    names = declared_symbols()
    assert len(names) > 50
    for n in names:
        assert hasattr(lib, n), "libsrslte_fec_b200.so does not export " + n


def test_no_oracle_or_torch_dependency(lib):
    """the product library links neither the oracle nor torch nor a dynamic CUDA runtime"""
    import subprocess
    out = subprocess.check_output(["ldd", b.LIB_PATH], text=True)
    assert "oracle" not in out and "torch" not in out and "libcudart" not in out
    syms = subprocess.check_output(["nm", "-D", "--defined-only", b.LIB_PATH], text=True)
    assert "orc_" not in syms


def test_segmentation_matches_oracle(lib, port):
    for tbs in list(range(16, 100000, 8 * 41)) + [75376, 97896, 6120, 40, 75000, 0]:
        rc, seg = b.cbsegm(tbs)
        rc2, seg2 = port.cbsegm(tbs)
        assert rc == rc2 and seg == seg2, tbs
    for i, K in enumerate(all_K()):
        assert lib.srslte_cbsegm_cbsize(C.c_uint32(i)) == K
        assert lib.srslte_cbsegm_cbindex(C.c_uint32(K)) == i
        assert lib.srslte_cbsegm_cbindex(C.c_uint32(K - 1)) == i
        assert lib.srslte_cbsegm_cbsize_isvalid(C.c_uint32(K)) and not lib.srslte_cbsegm_cbsize_isvalid(C.c_uint32(K + 1))
        assert lib.srslte_tdec_autoimp_get_subblocks(C.c_uint32(K)) == lanes16(K)
        assert lib.srslte_tdec_autoimp_get_subblocks_8bit(C.c_uint32(K)) == lanes8(K)
    assert lib.srslte_cbsegm_cbsize(C.c_uint32(188)) == -1 and lib.srslte_cbsegm_cbindex(C.c_uint32(6145)) == -1


def test_host_tables_match_oracle(lib, port):
    for K in all_K():
        for lanes in (1, 8, 16, 32):
            if lanes > 1 and K % lanes:
                continue
            f, r = b.qpp_table(K, lanes)
            f2, r2 = port.qpp(K, lanes)
            assert (f == f2).all() and (r == r2).all(), (K, lanes)
        for rv in range(4):
            for lanes in {0, lanes16(K), lanes8(K)}:
                assert (b.rm_table(K, rv, lanes) == port.rm_table(K, rv, lanes)).all(), (K, rv, lanes)


def test_crc_init_table(lib):
    c = b.Crc()
    assert lib.srslte_crc_init(C.byref(c), C.c_uint32(b.CRC24A), C.c_int(24)) == 0
    assert c.order == 24 and c.crcmask == 0xFFFFFF and c.table[1] == (b.CRC24A & 0xFFFFFF)
    assert lib.srslte_crc_init(C.byref(c), C.c_uint32(b.CRC24A), C.c_int(23)) == -1


@pytest.mark.skipif(os.path.exists("/dev/nvidiactl"), reason="a GPU is present")
def test_fails_loudly_without_gpu(lib):
    with pytest.raises(b.B200Error) as ei:
        b.Context(0)
    assert "no CUDA device" in str(ei.value) or "-3" in str(ei.value)
    td = b.Tdec()
    assert lib.srslte_tdec_init(C.byref(td), C.c_uint32(6144)) == -1


def _build_c_example(tmp_path):
    exe = str(tmp_path / "c_host_example")
    libdir = os.path.join(ROOT, "srsran_b200")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "c_host_example.c"),
                           "-L" + libdir, "-lsrslte_fec_b200", "-Wl,-rpath," + libdir, "-o", exe])
    return exe


def test_c_host_example_compiles_as_c99_and_links(tmp_path):
    """the headers are plain C (no C++-isms, no CUDA types) and a C program links against the shared library: the
    reference's host code is C.  Without a GPU the program stops at context creation with the library's message."""
    import srsran_b200 as b
    b.lib()  # make sure the library is built
    exe = _build_c_example(tmp_path)
    import torch
    if not torch.cuda.is_available():
        p = subprocess.run([exe], capture_output=True, text=True, timeout=60)
        assert p.returncode == 1 and "no CPU fallback" in p.stderr


@pytest.mark.gpu
def test_c_host_example_runs(tmp_path):
    """examples/c_host_example.c on the device: transmit mirror -> batched decode_tb -> the decode_tb_cb loop of sch.c:363-488
    written with the drop-in srslte_* symbols"""
    exe = _build_c_example(tmp_path)
    p = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert p.returncode == 0, p.stdout + p.stderr
    assert "batched decode_tb: ok" in p.stdout and "decode_tb_cb loop with srslte_* symbols: ok" in p.stdout


def test_reference_call_sites_compile_against_fec_h():
    """The reference's UNMODIFIED sch.c and pssch.c (every caller of the replaced symbols) type-check against
    include/srslte_b200/fec.h: tests/shim_include replaces srslte/phy/fec/{crc,turbodecoder,softbuffer,cbsegm}.h, everything
    else comes from the reference's own include tree.  Implicit declarations, incompatible pointers and int conversions are
    errors, so a missing symbol or a changed signature fails here."""
    ref = os.environ.get("SRSLTE_REFERENCE", "/root/reference")
    if not os.path.isdir(os.path.join(ref, "lib", "src", "phy", "phch")):
        pytest.skip("reference tree not present")
    import subprocess
    import tempfile
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    with tempfile.TemporaryDirectory() as tmp:
        os.makedirs(os.path.join(tmp, "srslte"))
        src = open(os.path.join(ref, "lib", "include", "srslte", "version.h.in")).read()
        for k, v in (("MAJOR", "20"), ("MINOR", "10"), ("PATCH", "1"), ("STRING", "20.10.1")):
            src = src.replace("@SRSLTE_VERSION_%s@" % k, v)
        open(os.path.join(tmp, "srslte", "version.h"), "w").write(src)
        for f in ("sch.c", "pssch.c"):
            cmd = ["gcc", "-std=gnu99", "-fsyntax-only", "-w", "-Werror=implicit-function-declaration", "-Werror=incompatible-pointer-types",
                   "-Werror=int-conversion", "-mavx2", "-mfma", "-DLV_HAVE_SSE", "-DLV_HAVE_AVX", "-DLV_HAVE_AVX2",
                   "-I" + os.path.join(root, "tests", "shim_include"), "-I" + os.path.join(root, "include"), "-I" + tmp,
                   "-I" + os.path.join(ref, "lib", "include"), os.path.join(ref, "lib", "src", "phy", "phch", f)]
            r = subprocess.run(cmd, capture_output=True, text=True)
            assert r.returncode == 0, r.stderr[-2000:]
