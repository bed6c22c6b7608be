"""TEST INFRASTRUCTURE ONLY: builds tests/emu/libfft_emu.so = the CUDA kernel SOURCE compiled for the
host through tests/emu/cuda_emu.h (see that header).  Used by tests/test_emu_kernels.py only."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.abspath(os.path.join(HERE, "..", "..", "chowdsp_fft_b200", "csrc"))
# CFB_EMU_DEFINES="-DX=1 -DY=0": build (and load) a variant of the kernel source with A/B switches set
DEFINES = os.environ.get("CFB_EMU_DEFINES", "").split()
SO = os.path.join(HERE, "libfft_emu.so" if not DEFINES else "libfft_emu_ab.so")


def build(force: bool = False) -> str:
    srcs = [os.path.join(HERE, "emu_driver.cpp"), os.path.join(HERE, "cuda_emu.h"),
            os.path.join(CSRC, "fft_kernels.cuh"), os.path.join(CSRC, "elementwise_kernels.cuh"),
            os.path.join(CSRC, "pconv_kernel.cuh"), os.path.join(CSRC, "pipe_kernels.cuh"), os.path.join(CSRC, "mixed_kernels.cuh"), os.path.join(CSRC, "mixq_kernels.cuh"), os.path.join(CSRC, "tma.cuh"), os.path.join(CSRC, "large_kernels.cuh"), os.path.join(CSRC, "large_plan.h"), os.path.join(CSRC, "cluster_kernels.cuh")]
    if not force and not DEFINES and os.path.exists(SO) and all(os.path.getmtime(SO) >= os.path.getmtime(s) for s in srcs):
        return SO
    cmd = ["g++", "-std=c++20", "-O1", "-fPIC", "-shared", "-pthread", *DEFINES, f"-I{HERE}", f"-I{CSRC}", srcs[0], "-o", SO]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("emu build failed:\n" + r.stderr[-4000:])
    return SO


if __name__ == "__main__":
    print(build(force=True))
