"""The drop-in boundary, checked without a GPU: the shared library loads, exports exactly what
include/*.h declares (and nothing from the reference's leaked helpers), the headers are valid C and
C++, and the no-device behaviour is loud (no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
INCLUDE = os.path.join(ROOT, "include")


def declared_functions():
    names = []
    for hdr in ("chowdsp_fft.h", "chowdsp_fft_b200.h"):
        text = open(os.path.join(INCLUDE, hdr)).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        names += re.findall(r"\b([a-z_0-9]+)\s*\([^;{]*\)\s*;", text)
    return sorted(set(names))


@pytest.fixture(scope="module")
def lib_path():
    from chowdsp_fft_b200 import _lib

    if not os.path.exists(_lib.LIB_PATH):
        _lib.build()
    return _lib.LIB_PATH


def test_header_declares_the_reference_surface():
    fns = declared_functions()
    for name in ("fft_bytes_required", "fft_new_setup", "fft_new_setup_preallocated", "fft_destroy_setup",
                 "fft_simd_width_bytes", "fft_transform", "fft_transform_unordered", "fft_convolve_unordered",
                 "fft_accumulate", "aligned_malloc", "aligned_free"):  # reference chowdsp_fft.h:81-163
        assert name in fns
    assert "fft_transform_batched" in fns and "fft_convolve_unordered_batched" in fns


def test_library_exports_every_declared_symbol(lib_path):
    from chowdsp_fft_b200 import _lib

    L = C.CDLL(lib_path)
    for name in declared_functions():
        assert hasattr(L, name), f"{name} declared in include/ but not exported"
    assert sorted(_lib.EXPORTED) == declared_functions()
    out = subprocess.run(["nm", "-D", "--defined-only", lib_path], capture_output=True, text=True).stdout
    exported = sorted(l.split()[-1] for l in out.splitlines() if " T " in l)
    assert exported == declared_functions(), "library exports symbols the headers do not declare"


@pytest.mark.parametrize("lang", ["c", "c++"])
def test_headers_compile_and_link(tmp_path, lib_path, lang):
    src = tmp_path / ("t.c" if lang == "c" else "t.cpp")
    ns = "" if lang == "c" else "using namespace chowdsp::fft;\n"
    src.write_text('#include "chowdsp_fft_b200.h"\n#include <stdio.h>\n' + ns + """
int main(void) {
    void* s = fft_new_setup(100, FFT_COMPLEX, true);   /* unsupported size -> NULL, never (void*)1 */
    if (s != NULL) return 1;
    float* p = (float*) aligned_malloc(1024);
    if (p == NULL || ((size_t) p & 63) != 0) return 2;
    p[0] = 1.f; aligned_free(p);
    if (fft_bytes_required(4096, FFT_COMPLEX, true) < 8 * 4096 + 96) return 3;  /* >= reference figure */
    if (fft_bytes_required(2048, FFT_REAL, true) < 4 * 2048 + 96) return 4;
    printf("launches=%llu err=%s\\n", fft_b200_launch_count(), fft_b200_last_error());
    return 0;
}
""")
    exe = tmp_path / "t"
    cc = ["gcc", "-std=c11"] if lang == "c" else ["g++", "-std=c++17"]
    r = subprocess.run(cc + [str(src), f"-I{INCLUDE}", lib_path, f"-Wl,-rpath,{os.path.dirname(lib_path)}", "-o", str(exe)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([str(exe)], capture_output=True, text=True, env=dict(os.environ, CHOWDSP_FFT_B200_QUIET="1"))
    assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)
    assert "unsupported FFT size" in r.stdout


def test_no_device_means_loud_failure_not_a_cpu_path(lib_path):
    import chowdsp_fft_b200 as cf

    if cf.device_available():
        pytest.skip("a CUDA device is present")
    os.environ["CHOWDSP_FFT_B200_QUIET"] = "1"
    with pytest.raises(cf.FFTError, match="no CPU fallback"):
        cf.fft_new_setup(1024, cf.FFT_REAL)
    with pytest.raises(cf.FFTError, match="unsupported FFT size"):
        cf.fft_new_setup(224, cf.FFT_REAL)  # 2^5 * 7: only the factors 2, 3, 5 are supported (as in the reference)
    assert cf.launch_count() == 0


def test_product_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under chowdsp_fft_b200/ may reference it."""
    pkg = os.path.join(ROOT, "chowdsp_fft_b200")
    for dirpath, _, files in os.walk(pkg):
        if os.sep + "build" in dirpath:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")) or f == "Makefile":
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert "oracle" not in text.lower() or f == "_lib.py" and False, f"{f} mentions the oracle"
                assert "cuda_emu" not in text or f == "fft_kernels.cuh", f"{f} references the test shim"
