"""The C-ABI library loads and exports every symbol include/*.h declares (no compute calls: this
runs without a GPU).  Also checks that the four reference symbols (csrc/src/ffi.rs:4-101) are there
under their exact names, and that the static archive the Rust build links exists."""
import ctypes as C
import re
import subprocess
from pathlib import Path

import pytest

from atoma_infer_b200 import _lib

ROOT = Path(__file__).resolve().parent.parent
HEADERS = sorted((ROOT / "include").glob("*.h"))
REFERENCE_SYMBOLS = ["run_mha", "copy_blocks_f16", "copy_blocks_bf16", "reshape_and_cache_flash"]


def declared_functions():
    names = []
    for h in HEADERS:
        src = re.sub(r"/\*.*?\*/", "", h.read_text(), flags=re.S)
        src = re.sub(r"//[^\n]*", "", src)
        src = re.sub(r"^\s*#[^\n]*", "", src, flags=re.M)
        src = re.sub(r"typedef\s+struct\s+\w+\s*\{.*?\}\s*\w+\s*;", "", src, flags=re.S)
        for m in re.finditer(r"\b([A-Za-z_]\w*)\s*\(([^;{}]*)\)\s*;", src):
            names.append(m.group(1))
    return sorted(set(names))


def test_headers_declare_the_reference_abi():
    names = declared_functions()
    for s in REFERENCE_SYMBOLS:
        assert s in names


def test_library_exports_every_declared_symbol():
    lib = C.CDLL(str(_lib.LIB_PATH))
    names = declared_functions()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/ but not exported by {_lib.LIB_PATH.name}"


def test_python_binding_covers_every_declared_symbol():
    assert sorted(_lib.SIGNATURES) == declared_functions()


def test_run_mha_signature_has_the_reference_arity():
    # csrc/src/ffi.rs:4-66 lists 47 parameters
    assert len(_lib.SIGNATURES["run_mha"][1]) == 47
    assert len(_lib.SIGNATURES["run_mha_stream"][1]) == 48
    assert len(_lib.SIGNATURES["reshape_and_cache_flash"][1]) == 14
    assert len(_lib.SIGNATURES["copy_blocks_bf16"][1]) == 7


def test_static_archive_for_the_rust_link_line():
    # csrc/build.rs:104-113 links `static=flashattention`
    a = _lib.LIB_PATH.with_name("libflashattention.a")
    assert a.exists()
    out = subprocess.run(["nm", "-g", "--defined-only", str(a)], capture_output=True, text=True).stdout
    for s in REFERENCE_SYMBOLS:
        assert re.search(rf"\bT {s}\b", out), s


def test_version_and_status_calls_work_without_a_gpu():
    L = _lib.lib()
    assert b"sm_100a" in L.atoma_version()
    assert L.atoma_last_status() == 0
    assert L.atoma_launch_count() >= 0


def test_sass_is_blackwell_native():
    """Built for sm_100a, and the SASS carries what B200_PROFILING.md lists as proof of a Blackwell-native
    kernel: TMA tensor loads (UTMALDG), bulk async copies (UBLKCP), tcgen05.mma (UTCHMMA) with accumulators
    read back from tensor memory (LDTM) — per kernel family, not just somewhere in the library."""
    out = subprocess.run(["cuobjdump", "-lelf", str(_lib.LIB_PATH)], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    sass = subprocess.run(["cuobjdump", "-sass", str(_lib.LIB_PATH)], capture_output=True, text=True).stdout
    per_kernel, name = {}, None
    for line in sass.splitlines():
        if "Function :" in line:
            name = line.split("Function :")[1].strip()
            per_kernel[name] = set()
        elif name is not None:
            for op in ("UTMALDG", "UBLKCP", "UTCHMMA", "LDTM", "STTM", "HMMA", "UTCBAR"):
                if op in line:
                    per_kernel[name].add(op)
    fam = lambda key: set().union(*[ops for n, ops in per_kernel.items() if key in n])
    assert {"UTMALDG", "UTCHMMA", "LDTM", "UTCBAR"} <= fam("gemm_skinny_kernel"), fam("gemm_skinny_kernel")
    assert {"UTMALDG", "UTCHMMA", "LDTM", "UTCBAR"} <= fam("gemm_chain2_kernel"), fam("gemm_chain2_kernel")
    assert {"UTMALDG", "UTCHMMA", "LDTM"} <= fam("gemm_tn_stream_kernel")
    assert {"UTMALDG", "UTCHMMA", "LDTM", "STTM"} <= fam("prefill_attn_kernel")
    assert {"UTMALDG", "HMMA"} <= fam("decode_paged_kernel")      # GQA rows cannot fill a 128-row tcgen05 tile
    assert "UBLKCP" in fam("copy_blocks_bulk_kernel")


def test_library_matches_the_source_tree():
    """atoma_source_hash() (compiled in by csrc/Makefile) == the hash of the sources in the tree: the binary the
    tests load was built from exactly these files."""
    L = _lib.lib()
    assert L.atoma_source_hash().decode() == _lib.source_hash()
    assert L.atoma_version().decode().endswith("src:" + _lib.source_hash())


def test_ctypes_struct_mirrors_have_the_library_layout():
    """include/atoma_llama.h's structs as the library was compiled against the ctypes mirrors in _lib.py (the
    vocab-parallel fields were appended to atoma_llama_model in round 2: a stale mirror would silently misread)."""
    lib = _lib.load()
    assert lib.atoma_llama_abi_sizeof(0) == C.sizeof(_lib.LlamaLayer)
    assert lib.atoma_llama_abi_sizeof(1) == C.sizeof(_lib.LlamaModel)
    assert lib.atoma_llama_abi_sizeof(2) == C.sizeof(_lib.DecodeBatch)
    assert lib.atoma_llama_abi_sizeof(3) == -1


@pytest.mark.parametrize("header", [h.name for h in HEADERS])
def test_headers_are_plain_c99(header, tmp_path):
    """The boundary is a C ABI: every header compiles as C99 on its own (what cgo / bindgen / a C host would see) —
    no C++ types, no torch types, no missing includes."""
    import shutil
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    src = tmp_path / "t.c"
    src.write_text(f'#include "{header}"\nint main(void) {{ return 0; }}\n')
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", f"-I{ROOT / 'include'}", "-fsyntax-only", str(src)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_static_archive_links_with_the_reference_link_line(tmp_path):
    """csrc/build.rs:104-113 links `static=flashattention`, `dylib=cudart`, `dylib=stdc++` and nothing else: a C program
    that references symbols of all three headers links against libflashattention.a with exactly those libraries and runs
    (version / status calls only — no GPU needed)."""
    import os
    import shutil
    cudart = Path("/usr/local/cuda/lib64")
    if shutil.which("gcc") is None or not (cudart / "libcudart.so").exists():
        pytest.skip("needs gcc and the CUDA runtime library")
    src = tmp_path / "link_check.c"
    src.write_text('#include "atoma_flash.h"\n#include "atoma_llama.h"\n#include "atoma_ops.h"\n#include <stdio.h>\n'
                   'int main(void) {\n'
                   '    void *keep[] = {(void *)run_mha, (void *)copy_blocks_bf16, (void *)reshape_and_cache_flash,\n'
                   '                    (void *)atoma_llama_decode_step, (void *)atoma_prepare_input};\n'
                   '    printf("%s %lld %d\\n", atoma_version(), (long long)atoma_llama_abi_sizeof(1), (int)(sizeof keep / sizeof keep[0]));\n'
                   '    return atoma_last_status();\n}\n')
    exe = tmp_path / "link_check"
    lib_dir = _lib.LIB_PATH.parent
    r = subprocess.run(["gcc", "-std=c99", f"-I{ROOT / 'include'}", str(src), f"-L{lib_dir}", "-l:libflashattention.a",
                        f"-L{cudart}", "-lcudart", "-lstdc++", "-lm", "-ldl", "-lpthread", "-o", str(exe)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    env = dict(os.environ, LD_LIBRARY_PATH=f"{cudart}:{os.environ.get('LD_LIBRARY_PATH', '')}")
    out = subprocess.run([str(exe)], capture_output=True, text=True, env=env)
    assert out.returncode == 0 and "sm_100a" in out.stdout, out.stdout + out.stderr
