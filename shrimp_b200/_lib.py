"""ctypes binding of libshrimp_b200.so (the C ABI declared in include/shrimp_b200.h).

The library is built in-tree by ``__graft_entry__.build()`` / ``make -C shrimp_b200/csrc``.  There
is no fallback: if the shared object is missing, or no CUDA device is usable, the calls raise.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libshrimp_b200.so")

_lib = None


class ShrimpGpuError(RuntimeError):
    pass


class SwParams(C.Structure):
    _fields_ = [(n, C.c_int) for n in (
        "match", "mismatch", "a_gap_open", "a_gap_ext", "b_gap_open", "b_gap_ext", "crossover",
        "use_colours", "anchor_width", "indel_taboo_len", "max_read_len", "max_window_len")]


class MapParamsC(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("window_len", "window_overlap", "window_gen_threshold", "sw_vect_threshold",
                                          "sw_full_threshold", "score_alpha", "score_beta")] + \
               [(n, C.c_int32) for n in ("match_mode", "num_outputs", "num_tmp_outputs", "gapless", "hash_filter_calls",
                                         "use_regions", "region_bits", "region_overlap", "Gflag", "Tflag", "strata",
                                         "max_alignments", "compute_mapping_qualities")] + \
               [("list_cutoff", C.c_uint32), ("crossover_scores", C.c_void_p), ("crossover_stride", C.c_int32),
                ("read_quals", C.c_void_p), ("qual_stride", C.c_int32), ("qual_delta", C.c_int32),
                ("qual_vector_offset", C.c_int32), ("use_sanger_qvs", C.c_int32), ("pr_xover", C.c_double)]


class HitC(C.Structure):
    _fields_ = [("read_idx", C.c_int32), ("cn", C.c_int32), ("gen_st", C.c_int32), ("w_len", C.c_int32),
                ("g_off", C.c_int64),
                ("score_vector", C.c_int32), ("score_full", C.c_int32), ("pass2_key", C.c_int32),
                ("score_max", C.c_int32), ("matches", C.c_int32), ("sw_score", C.c_int32),
                ("posterior", C.c_double),
                ("read_start", C.c_int32), ("rmapped", C.c_int32), ("genome_start", C.c_int32), ("gmapped", C.c_int32),
                ("sfr_matches", C.c_int32), ("mismatches", C.c_int32), ("insertions", C.c_int32),
                ("deletions", C.c_int32), ("crossovers", C.c_int32), ("edit_len", C.c_int32), ("edit_off", C.c_int64),
                ("hit_slot", C.c_int32), ("st", C.c_int32), ("score_window_gen", C.c_int32), ("reserved", C.c_int32)]


class StageHitC(C.Structure):
    _fields_ = [("read_idx", C.c_int32), ("st", C.c_int32), ("cn", C.c_int32), ("w_len", C.c_int32),
                ("g_off", C.c_int64),
                ("score_window_gen", C.c_int32), ("matches", C.c_int32), ("score_max", C.c_int32),
                ("score_vector", C.c_int32), ("pct_score_vector", C.c_int32),
                ("ax", C.c_int32), ("ay", C.c_int32), ("alen", C.c_int32), ("awidth", C.c_int32)]


class FullTaskC(C.Structure):
    _fields_ = [("goff", C.c_uint32)] + [(n, C.c_int32) for n in (
        "glen", "read_idx", "rlen", "threshscore", "maxscore", "revcmpl", "ax", "ay", "alen", "awidth", "initbp")]


class FullResultC(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("score", "read_start", "rmapped", "genome_start", "gmapped", "matches",
                                         "mismatches", "insertions", "deletions", "crossovers", "edit_len")] + \
               [("edit_off", C.c_int64)]


class PairParamsC(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("pair_mode", "min_insert_size", "max_insert_size", "half_paired")]


class PairC(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("pair_idx", "score", "score_max", "key", "insert_size")] + \
               [("hit_idx", C.c_int32 * 2)]


class MapStatsC(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("list_entries", "surviving_entries", "anchors", "hits", "heap_replays",
                                          "vector_tasks", "vector_calls", "vector_cells", "vector_bypassed",
                                          "full_calls", "full_cells", "device_vector_cells", "scan_big_strands",
                                          "scan_global_strands", "post_sw_columns")]


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ShrimpGpuError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU fallback for the gmapper hot path)")
    L = C.CDLL(LIB_PATH)
    vp, i32, u64 = C.c_void_p, C.c_int, C.c_uint64
    L.shrimp_gpu_device_count.restype = i32
    L.shrimp_gpu_create.argtypes = [i32, C.POINTER(vp)]
    L.shrimp_gpu_create.restype = i32
    L.shrimp_gpu_destroy.argtypes = [vp]
    L.shrimp_gpu_destroy.restype = None
    L.shrimp_gpu_last_error.restype = C.c_char_p
    L.shrimp_gpu_launch_count.argtypes = [vp]
    L.shrimp_gpu_launch_count.restype = u64
    L.shrimp_gpu_stage_times.argtypes = [vp, C.POINTER(C.c_char_p), C.POINTER(C.c_float), C.POINTER(u64), i32]
    L.shrimp_gpu_stage_times.restype = i32
    L.shrimp_gpu_stage_times_reset.argtypes = [vp]
    L.shrimp_gpu_stage_times_reset.restype = None
    L.shrimp_gpu_sw_setup.argtypes = [vp, C.POINTER(SwParams)]
    L.shrimp_gpu_sw_setup.restype = i32
    L.shrimp_gpu_sw_vector_batch.argtypes = [vp, vp, C.c_size_t, vp, vp, i32, i32, i32, vp, vp, vp, vp, vp, vp]
    L.shrimp_gpu_sw_vector_batch.restype = i32
    L.shrimp_gpu_sw_gapless_batch.argtypes = [vp, vp, C.c_size_t, vp, vp, i32, i32, i32, vp, vp, vp, vp, vp, vp, vp, vp]
    L.shrimp_gpu_sw_gapless_batch.restype = i32
    L.shrimp_gpu_sw_full_batch.argtypes = [vp, vp, C.c_size_t, vp, i32, i32, i32, vp, i32, vp, vp, C.c_int64,
                                           C.POINTER(C.c_int64)]
    L.shrimp_gpu_sw_full_batch.restype = i32
    L.shrimp_gpu_sw_full_batch_xover.argtypes = [vp, vp, C.c_size_t, vp, i32, i32, i32, vp, i32, vp, i32, vp, vp,
                                                 C.c_int64, C.POINTER(C.c_int64)]
    L.shrimp_gpu_sw_full_batch_xover.restype = i32
    L.shrimp_gpu_dpx_peak.argtypes = [vp, C.POINTER(C.c_double)]
    L.shrimp_gpu_dpx_peak.restype = i32
    L.shrimp_gpu_fp64_peak.argtypes = [vp, C.POINTER(C.c_double)]
    L.shrimp_gpu_fp64_peak.restype = i32
    L.shrimp_gpu_set_host_threads.argtypes = [C.c_int]
    L.shrimp_gpu_set_host_threads.restype = i32
    L.shrimp_gpu_glibc_explog.argtypes = [vp, vp, C.c_int, vp, vp]
    L.shrimp_gpu_glibc_explog.restype = i32
    L.shrimp_gpu_hash_windows.argtypes = [vp, vp, u64, vp, vp, C.c_int, vp]
    L.shrimp_gpu_hash_windows.restype = i32
    L.shrimp_gpu_genome_load.argtypes = [vp, i32, vp, vp, i32]
    L.shrimp_gpu_genome_load.restype = i32
    L.shrimp_gpu_share_genome.argtypes = [vp, vp]
    L.shrimp_gpu_share_genome.restype = i32
    L.shrimp_gpu_genome_export.argtypes = [vp, i32, vp, C.c_size_t]
    L.shrimp_gpu_genome_export.restype = i32
    L.shrimp_gpu_index_build.argtypes = [vp, i32, vp, vp, vp, i32]
    L.shrimp_gpu_index_build.restype = i32
    L.shrimp_gpu_index_nbuckets.argtypes = [vp, i32, C.POINTER(C.c_uint32), C.POINTER(u64)]
    L.shrimp_gpu_index_nbuckets.restype = i32
    L.shrimp_gpu_index_export.argtypes = [vp, i32, vp, vp, C.POINTER(u64)]
    L.shrimp_gpu_index_export.restype = i32
    L.shrimp_gpu_projection_save.argtypes = [vp, C.c_char_p, vp]
    L.shrimp_gpu_projection_save.restype = i32
    L.shrimp_gpu_projection_load.argtypes = [vp, C.c_char_p]
    L.shrimp_gpu_projection_load.restype = i32
    L.shrimp_gpu_num_contigs.argtypes = [vp]
    L.shrimp_gpu_num_contigs.restype = i32
    L.shrimp_gpu_contig_name.argtypes = [vp, i32]
    L.shrimp_gpu_contig_name.restype = C.c_char_p
    L.shrimp_gpu_map_reads.argtypes = [vp, C.POINTER(MapParamsC), i32, vp, i32, vp, vp, vp, C.c_int64, vp, vp,
                                       C.c_int64, C.POINTER(C.c_int64), C.POINTER(C.c_int64), vp, C.c_int64,
                                       C.POINTER(C.c_int64), C.POINTER(MapStatsC)]
    L.shrimp_gpu_map_reads.restype = i32
    L.shrimp_gpu_map_pairs.argtypes = [vp, C.POINTER(MapParamsC), C.POINTER(PairParamsC), i32, vp, i32, vp, vp,
                                       vp, C.c_int64, C.POINTER(C.c_int64), vp, C.c_int64, C.POINTER(C.c_int64),
                                       vp, vp, vp, C.c_int64, C.POINTER(C.c_int64), C.POINTER(MapStatsC)]
    L.shrimp_gpu_map_pairs.restype = i32
    L.shrimp_gpu_map_pairs_resident.argtypes = [vp, C.POINTER(MapParamsC), C.POINTER(PairParamsC), C.POINTER(MapStatsC)]
    L.shrimp_gpu_map_pairs_resident.restype = i32
    L.shrimp_gpu_map_resident.argtypes = [vp, C.POINTER(MapParamsC), C.POINTER(MapStatsC)]
    L.shrimp_gpu_map_resident.restype = i32
    L.shrimp_gpu_last_transfer_bytes.argtypes = [vp, C.POINTER(u64), C.POINTER(u64)]
    L.shrimp_gpu_last_transfer_bytes.restype = i32
    L.shrimp_gpu_event_record.argtypes = [vp, i32]
    L.shrimp_gpu_event_record.restype = i32
    L.shrimp_gpu_event_elapsed_ms.argtypes = [vp, C.POINTER(C.c_float)]
    L.shrimp_gpu_event_elapsed_ms.restype = i32
    L.shrimp_gpu_flush_l2.argtypes = [vp]
    L.shrimp_gpu_flush_l2.restype = i32
    _lib = L
    return L


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().shrimp_gpu_last_error().decode(errors="replace")
        raise ShrimpGpuError(f"{what} failed (rc={rc}): {msg}")
