"""ctypes binding of csrc/libmcgpu.so (the C ABI declared in include/mc_b200.h)."""
from __future__ import annotations

import ctypes as C
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
# MC_B200_LIB: another BUILD of the same library (kernel A/B experiments, profiles/scripts/build_variant.sh); never a fallback
_LIB = os.environ.get("MC_B200_LIB") or os.path.join(_HERE, "csrc", "libmcgpu.so")
_HEADER = os.path.join(os.path.dirname(_HERE), "include", "mc_b200.h")


class McError(RuntimeError):
    """A failing libmcgpu call; `.code` is the mc_status, the text follows the reference's exception messages."""

    def __init__(self, code, msg):
        super().__init__(f"[mc_status {code}] {msg}")
        self.code = code
        self.msg = msg


def lib_path():
    return _LIB


_lib = None

vp, u64, i64, sz, i32 = C.c_void_p, C.c_uint64, C.c_int64, C.c_size_t, C.c_int
cp = C.c_char_p

_SIGNATURES = {
    "mc_last_error": (cp, []),
    "mc_version": (cp, []),
    "mc_ctx_create": (i32, [i32, C.POINTER(vp)]),
    "mc_ctx_destroy": (None, [vp]),
    "mc_ctx_set_chunk_bytes": (i32, [vp, sz]),
    "mc_ctx_launches": (u64, [vp]),
    "mc_ctx_trim": (i32, [vp]),
    "mc_ctx_arena_stats": (i32, [vp, C.POINTER(u64), C.POINTER(u64)]),
    "mc_table_create": (i32, [vp, i32, u64, i32, i32, C.POINTER(vp)]),
    "mc_table_free": (None, [vp]),
    "mc_count_files": (i32, [vp, C.POINTER(cp), i32, i32, C.POINTER(vp)]),
    "mc_count_file": (i32, [vp, cp, vp]),
    "mc_count_bytes": (i32, [vp, vp, sz, i32, vp]),
    "mc_count_device": (i32, [vp, vp, sz, i32, vp]),
    "mc_table_size": (u64, [vp]),
    "mc_table_stats": (i32, [vp, C.POINTER(u64), C.POINTER(u64), C.POINTER(u64), C.POINTER(u64)]),
    "mc_table_add_pairs": (i32, [vp, vp, vp, sz]),
    "mc_table_get": (i32, [vp, vp, sz, vp]),
    "mc_table_dump": (i32, [vp, i32, vp, vp, sz, C.POINTER(sz)]),
    "mc_ctx_host_ms": (i32, [vp, vp]),
    "mc_classify_bytes": (i32, [vp, vp, vp, sz, i32, C.c_double, C.c_double, i32, vp, sz, C.POINTER(sz)]),
    "mc_java_format_6_2f": (i32, [C.c_float, cp]),
    "mc_table_dump_device": (i32, [vp, i32, vp, vp, sz, C.POINTER(sz)]),
    "mc_table_add_pairs_device": (i32, [vp, vp, vp, sz]),
    "mc_table_filter": (i32, [vp, vp, i32, C.POINTER(vp)]),
    "mc_env_find": (i32, [vp, vp, C.POINTER(cp), C.POINTER(cp), i32, i32, i32, i64, i64, i32, i32, cp, C.POINTER(i32)]),
    "mc_env_find_ex": (i32, [vp, vp, C.POINTER(cp), C.POINTER(cp), i32, C.POINTER(cp), i32, i32, i32, i64, i64, i32, i32, i32, i32, cp, C.POINTER(i32)]),
    "mc_env_extend_counts": (i32, [vp, vp, i32, C.POINTER(i32)]),
    "mc_bfs": (i32, [vp, vp, cp, i32, i32, i64, i64, i32, vp, vp, vp, sz, C.POINTER(sz)]),
    "mc_env_kmers": (i32, [vp, vp, cp, i32, i32, i64, i64, i32, vp, vp, sz, C.POINTER(sz)]),
    "mc_env_from_graph_txt": (i32, [vp, cp, cp, i32, i32, cp]),
    "mc_targets_load": (i32, [cp, C.POINTER(vp)]),
    "mc_targets_count": (i32, [vp]),
    "mc_targets_seq": (cp, [vp, i32]),
    "mc_targets_name": (cp, [vp, i32]),
    "mc_targets_free": (None, [vp]),
    "mc_env_multi": (i32, [vp, C.POINTER(cp), i32, cp, i32, cp]),
    "mc_table_print_kmers": (i32, [vp, i32, cp, cp, C.POINTER(u64)]),
    "mc_table_load_kmers": (i32, [vp, cp, i32]),
    "mc_extract_partition_device": (i32, [vp, vp, sz, i32, i32, i32, vp, C.POINTER(u64), vp]),
    "mc_max_kmers": (u64, [sz, i32, i32]),
    "mc_table_insert_hashed_device": (i32, [vp, vp, sz]),
    "mc_xchg_create": (i32, [vp, i32, i32, u64, C.POINTER(vp)]),
    "mc_xchg_handle": (i32, [vp, vp]),
    "mc_xchg_open": (i32, [vp, i32, vp]),
    "mc_xchg_scatter_device": (i32, [vp, vp, sz, i32, i32, i32, vp]),
    "mc_xchg_insert": (i32, [vp, i32, vp, C.POINTER(u64)]),
    "mc_xchg_dump_owners": (i32, [vp, vp, i32, C.POINTER(u64)]),
    "mc_xchg_flush_and_dump": (i32, [vp, vp, i32, C.POINTER(u64), C.POINTER(i32)]),
    "mc_xchg_merge": (i32, [vp, i32, vp, C.POINTER(u64)]),
    "mc_table_copy_stats": (i32, [vp, vp]),
    "mc_xchg_destroy": (None, [vp]),
    "mc_synth_generate_device": (i32, [vp, vp, sz, u64, u64, vp]),
    "mc_device_alloc": (i32, [vp, sz, C.POINTER(vp)]),
    "mc_device_free": (i32, [vp, vp]),
    "mc_memcpy_d2h": (i32, [vp, vp, vp, sz]),
    "mc_memcpy_h2d": (i32, [vp, vp, vp, sz]),
    "mc_ctx_sync": (i32, [vp]),
    "mc_event_record": (i32, [vp, i32]),
    "mc_event_elapsed_ms": (i32, [vp, i32, i32, C.POINTER(C.c_float)]),
    "mc_ctx_stage_ms": (i32, [vp, C.POINTER(C.c_float)]),
    "mc_ctx_stage_ms_all": (i32, [vp, C.POINTER(C.c_float)]),
    "mc_ctx_stage_reset": (i32, [vp]),
    "mc_ctx_set_tuning": (i32, [vp, cp, C.c_longlong]),
    "mc_table_flush_pending": (i32, [vp]),
    "mc_table_expect_keys": (i32, [vp, u64]),
    "mc_table_checksum": (i32, [vp, i32, C.POINTER(u64)]),
    "mc_kmer_hashmap_order": (i32, [vp, sz, i32, sz, vp, C.POINTER(sz)]),
    "mc_detect_file_format": (i32, [cp, C.POINTER(i32)]),
    "mc_sniff_quality_offset": (i32, [vp, sz]),
}


def declared_symbols():
    """Every function name include/mc_b200.h declares (used by the CPU test that checks the exports)."""
    text = open(_HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mc_[a-z0-9_]+)\s*\(", text)))


def load_library():
    """Loads libmcgpu.so.  Raises (never falls back) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB):
        raise McError(-1, f"{_LIB} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(make -C metacherchant_b200/csrc). There is no CPU fallback.")
    lib = C.CDLL(_LIB)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        lib = load_library()
        raise McError(rc, (lib.mc_last_error() or b"").decode("utf-8", "replace"))
