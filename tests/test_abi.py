"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol include/mc_b200.h declares,
and refuses to compute without a GPU (no CPU fallback)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "metacherchant_b200", "csrc", "metacherchant")


@pytest.fixture(scope="module")
def lib():
    from metacherchant_b200 import _ffi
    if not os.path.exists(_ffi.lib_path()):
        subprocess.run(["make", "-C", os.path.join(ROOT, "metacherchant_b200", "csrc"), "-j", "8"], check=True)
    return _ffi.load_library()


def test_library_exports_every_declared_symbol(lib):
    from metacherchant_b200 import _ffi
    declared = _ffi.declared_symbols()
    assert len(declared) >= 40
    missing = [s for s in declared if not hasattr(lib, s)]
    assert not missing, missing
    # and the Python binding covers the whole header
    unbound = [s for s in declared if s not in _ffi._SIGNATURES]
    assert not unbound, unbound


def test_header_cites_the_reference_interfaces():
    text = open(os.path.join(ROOT, "include", "mc_b200.h")).read()
    for cite in ("src/io/IOUtils.java:217", "Long2ShortHashMap.java:159", "src/tools/EnvironmentFinderMain.java:218",
                 "src/tools/EnvironmentFinderMultiMain.java:81", "src/algo/OneSequenceCalculator.java"):
        assert cite in text, cite


def test_no_gpu_means_loud_failure(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from metacherchant_b200 import Context, McError
    with pytest.raises(McError) as e:
        Context(0)
    assert "no CPU fallback" in str(e.value)


def test_product_never_touches_the_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py may import/link/execute anything under oracle/."""
    pkg = os.path.join(ROOT, "metacherchant_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")) or f == "Makefile":
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert "mc_oracle" not in text and "oracle_ffi" not in text and "pyref" not in text, os.path.join(dirpath, f)
                assert "from oracle" not in text and "import oracle" not in text, os.path.join(dirpath, f)


def run_cli(*args, cwd=None):
    return subprocess.run([CLI, *args], capture_output=True, text=True, cwd=cwd)


def test_cli_flag_surface(tmp_path, lib):
    assert os.path.exists(CLI)
    r = run_cli("--help", cwd=tmp_path)
    assert r.returncode == 0
    for flag in ("--k", "--reads", "--seq", "--output", "--maxkmers", "--maxradius", "--coverage", "--bothdirs", "--chunklength", "--work-dir",
                 "--available-processors", "--tool"):
        assert flag in r.stdout, flag
    r = run_cli("--seq", "x.fa", "--output", "o", "--maxkmers", "5", cwd=tmp_path)
    assert r.returncode == 1 and "Mandatory argument --k (-k) not set" in r.stderr
    r = run_cli("-k", "31", "--seq", "x.fa", "--output", "o", cwd=tmp_path)
    assert r.returncode == 1 and "At least one of --maxkmers and --maxradius parameters should be set" in r.stderr
    for bad in (["--forcehash", "True"], ["--hash", "fnv1a"]):
        r = run_cli("-k", "31", "--seq", "x.fa", "--output", "o", "--maxkmers", "5", *bad, cwd=tmp_path)
        assert r.returncode == 1 and "not supported by the GPU path" in r.stderr, bad
    r = run_cli("-k", "33", "--seq", "x.fa", "--output", "o", "--maxkmers", "5", cwd=tmp_path)
    assert r.returncode == 1 and "k > 31" in r.stderr
    r = run_cli("--tool", "environment-finder-multi", "--seq", "x.fa", "--output", "o", cwd=tmp_path)
    assert r.returncode == 1 and "Mandatory argument --env (-e) not set" in r.stderr
    r = run_cli("--tool", "reads-classifier", "-k", "21", "-i", "g.fastq", cwd=tmp_path)
    assert r.returncode == 1 and "Mandatory argument --read-files (-r) not set" in r.stderr
    r = run_cli("--tool", "assembler", cwd=tmp_path)
    assert r.returncode == 1 and "Unknown tool" in r.stderr
    r = run_cli("-k", "31", "--seq", "x.fa", "--output", "o", "--maxkmers", "5", "--bogus", "1", cwd=tmp_path)
    assert r.returncode == 1 and "Unrecognized option: --bogus" in r.stderr
    assert (tmp_path / "workDir" / "log").exists()          # Tool.updateFileLoggers: work-dir/log


def test_jni_shim_compiles(tmp_path):
    """jni/mc_jni.cpp is source-only in this image (no JDK): type-check it against include/mc_b200.h with a stand-in <jni.h> that
    declares the JNI types and the JNIEnv members it uses (tests/jni_stub/jni.h).  Compile only: nothing is linked or run."""
    import shutil
    import subprocess
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else shutil.which("g++")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([cxx, "-std=c++17", "-fsyntax-only", "-Wall", "-Wextra", "-DMC_HAVE_JNI=1", "-I", os.path.join(root, "tests", "jni_stub"),
                        "-I", os.path.join(root, "include"), os.path.join(root, "jni", "mc_jni.cpp")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    src = open(os.path.join(root, "jni", "mc_jni.cpp")).read()
    for fn in ("createContext", "loadReads", "size", "get", "findEnvironments", "findEnvironmentsEx", "multiEnvironment", "classifyReads", "freeTable", "destroyContext"):
        assert f"Java_io_NativeCounter_{fn}" in src
