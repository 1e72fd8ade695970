"""ctypes loader of libscone_b200.so (built in-tree by scone_b200/csrc/build.sh)."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class EngineError(RuntimeError):
    pass


def library_path():
    # SB_LIBRARY: an instrumented build of the same library (profiles/round_profile.py); never a different implementation
    return os.environ.get("SB_LIBRARY") or os.path.join(_HERE, "libscone_b200.so")


class CycleResult(C.Structure):
    """sb_cycle_result (include/scone_b200.h)."""
    _fields_ = [("n_start", C.c_int32), ("n_sites", C.c_int32),
                ("start_wgt", C.c_double), ("end_wgt", C.c_double),
                ("imp_prod", C.c_double), ("imp_abs", C.c_double), ("scatter_prod", C.c_double), ("ana_leak", C.c_double),
                ("k_analog", C.c_double), ("k_implicit", C.c_double), ("k_cum", C.c_double), ("k_cum_std", C.c_double),
                ("n_segments", C.c_int64), ("n_collisions", C.c_int64), ("n_scores", C.c_int64), ("error", C.c_int32), ("max_history_segments", C.c_int32), ("n_xs_terms", C.c_int64)]


def load_library():
    """Load the engine. Raises EngineError when the CUDA extension has not been built."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = library_path()
    if not os.path.exists(path):
        raise EngineError("scone_b200: %s is missing -- run __graft_entry__.build() (scone_b200/csrc/build.sh). "
                          "There is no CPU fallback." % path)
    L = C.CDLL(path)
    vp, i32, i64, u64, dbl = C.c_void_p, C.c_int, C.c_int64, C.c_uint64, C.c_double
    dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int32)
    sig = {
        # C ABI (include/scone_b200.h)
        "sb_create": (i32, [C.POINTER(vp), i32]), "sb_destroy": (None, [vp]), "sb_last_error": (C.c_char_p, [vp]),
        "sb_launch_count": (i64, [vp]),
        "sb_load_geometry": (i32, [vp, vp]), "sb_load_mg_data": (i32, [vp, vp]),
        "sb_define_tallies": (i32, [vp, i32, vp, i32, i32, dbl]), "sb_set_options": (i32, [vp, vp]),
        "sb_bank_upload": (i32, [vp, i32, dp, dp, dp, ip]), "sb_bank_download": (i32, [vp, i32, ip, dp, dp, dp, ip]),
        "sb_bank_upload_ce": (i32, [vp, i32, dp, dp, dp, dp]), "sb_bank_download_ce": (i32, [vp, i32, ip, dp, dp, dp, dp]),
        "sb_load_ce_model": (i32, [vp, vp]), "sb_ce_nuclide_info": (i32, [vp, i32, ip, ip, ip]), "sb_ce_nuclide_data": (i32, [vp, i32, dp, dp, ip]),
        "sb_bank_size": (i32, [vp]), "sb_source_generate": (i32, [vp, i32, u64, i32]),
        "sb_run_cycle": (i32, [vp, u64, i32, dbl, i32, C.POINTER(CycleResult)]),
        "sb_resample": (i32, [vp, i32, u64]),
        "sb_run_cycle_resample": (i32, [vp, u64, i32, dbl, i32, i32, u64, C.POINTER(CycleResult)]),
        "sb_cycle_begin": (i32, [vp, u64, i32, dbl, i32, vp, ip]), "sb_cycle_end": (i32, [vp, vp, C.POINTER(CycleResult)]),
        "sb_cycle_end_resample_ranked": (i32, [vp, dp, i32, u64, i32, i32, ip, ip, C.POINTER(CycleResult)]),
        "sb_resample_ranked": (i32, [vp, i32, u64, i32, i32, ip, ip]), "sb_site_buffer_bytes": (C.c_size_t, [i32]),
        "sb_bank_export": (i32, [vp, i32, vp, i32, vp]), "sb_bank_splice": (i32, [vp, i32, i32, i32, vp, i32, vp]),
        "sb_tally_size": (i64, [vp, i32]), "sb_tally_read": (i32, [vp, i32, dp, dp, ip]), "sb_tally_last_bins": (i32, [vp, i32, dp]),
        "sb_geom_query": (i32, [vp, i64, dp, dp, dp, ip, ip]), "sb_mg_query": (i32, [vp, i64, ip, ip, dp, dp]),
        "sb_timer_begin": (i32, [vp]), "sb_timer_end": (i32, [vp, dp]),
        "sb_profile_enable": (i32, [vp, i32]),
        "sb_profile_read": (i32, [vp, dp, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
        "sb_run_cycle_resample_host": (i32, [vp, i32, dp, dp, dp, ip, dp, u64, i32, dbl, i32, i32, u64, ip, dp, dp, dp, ip, dp, dp, C.POINTER(CycleResult)]),
        "sb_profile_peer_stages": (i32, [vp, dp, dp]),
        "sb_flush_l2": (i32, [vp, C.c_size_t]), "sb_pinned_alloc": (vp, [C.c_size_t]), "sb_pinned_free": (None, [vp]),
        "sb_load_ce_data": (i32, [vp, vp]), "sb_ce_union_size": (i32, [vp]), "sb_ce_union": (i32, [vp, dp, dp]),
        "sb_ce_lookup": (i32, [vp, i64, dp, ip, dp, dp, dp]), "sb_ce_lookup_device": (i32, [vp, i64, vp, vp, vp, vp, vp]), "sb_ce_lookup_sorted_device": (i32, [vp, i64, vp, vp, vp, vp, vp]),
        "sb_ce_nuclide_index": (i32, [vp, i32, i64, dp, ip]), "sb_ce_last_kernel_ms": (i32, [vp, dp]),
        "sb_ce_memory": (i32, [vp, C.POINTER(i64), C.POINTER(i64), ip]),
        "sb_rng_query": (i32, [i64, C.POINTER(C.c_uint64), C.POINTER(C.c_int64), C.POINTER(C.c_uint64), dp]),
        "sb_math_query": (i32, [i64, dp, dp, dp, dp]),
        "sb_fastmath_check": (i32, [i64, C.c_uint64, i32, C.POINTER(i64)]),
        # host driver (scone_b200/csrc/host/physics_package.cpp)
        "sbh_last_error": (C.c_char_p, [vp]),
        "sbh_eigen_create": (vp, [C.c_char_p, C.c_char_p, i32, i32, i32]), "sbh_eigen_destroy": (None, [vp]),
        "sbh_engine": (vp, [vp]), "sbh_model_dump": (i32, [vp, C.c_char_p]),
        "sbh_eigen_info": (i32, [vp] + [ip] * 7),
        "sbh_eigen_total_pop": (i32, [vp]), "sbh_eigen_rng_state": (u64, [vp]), "sbh_eigen_set_rng_state": (None, [vp, u64]), "sbh_eigen_keff0": (dbl, [vp]),
        "sbh_eigen_generate_initial_state": (i32, [vp]),
        "sbh_eigen_cycle": (i32, [vp, i32, dp, C.POINTER(CycleResult)]),
        "sbh_eigen_cycle_host_buffers": (i32, [vp, i32, dp, C.POINTER(CycleResult)]),
        "sbh_eigen_cycle_begin": (i32, [vp, i32, dbl, vp, ip]), "sbh_eigen_cycle_end": (i32, [vp, i32, vp, C.POINTER(CycleResult)]),
        "sbh_eigen_resample_ranked": (i32, [vp, ip, ip, dp]),
        "sbh_eigen_cycle_end_resample_ranked": (i32, [vp, i32, dp, ip, ip, dp, C.POINTER(CycleResult)]),
        "sbh_workshare": (i32, [i32, i32, i32, ip, ip]), "sbh_balance_plan": (i32, [i32, i32, i32, ip, ip]),
        "sbh_eigen_download_bank": (i32, [vp]), "sbh_eigen_upload_bank": (i32, [vp]),
        "sbh_eigen_is_ce": (i32, [vp]), "sbh_eigen_is_fixed": (i32, [vp]), "sbh_fixed_cycle": (i32, [vp, C.POINTER(CycleResult)]),
        "sb_set_fixed_source": (i32, [vp, i32, i32]), "sb_source_point": (i32, [vp, i32, u64, i32, vp]), "sbh_ce_info": (i32, [vp, ip, ip]),
        "sb_bank_brood": (i32, [vp, i32, ip]), "sb_set_file_source": (i32, [vp, i64, dp, i32]), "sb_source_file": (i32, [vp, i32, u64, i32]),
        "sbh_eigen_print_source": (i32, [vp, i32]), "sbh_eigen_print_source_mode": (i32, [vp]),
        "sb_source_material": (i32, [vp, i32, u64, i32, vp]), "sb_geometry_bounds": (i32, [vp, dp]),
        "sbh_dict_get": (i32, [C.c_char_p, i32, C.c_char_p, C.c_char, C.c_char_p, i32]),
        "sb_peer_create": (i32, [vp, i32, i32, i32, vp]), "sb_bank_capacity": (i32, [vp]), "sb_peer_attach": (i32, [vp, vp, ip]), "sb_peer_capacity": (i32, [vp]),
        "sb_peer_set_timeout": (i32, [vp, dbl]),
        "sb_run_cycle_ranked_peer": (i32, [vp, u64, i32, dbl, i32, i32, u64, ip, C.POINTER(CycleResult)]),
        "sbh_eigen_cycle_peer": (i32, [vp, i32, dp, ip, C.POINTER(CycleResult)]),
        "sbh_ce_card_process": (i32, [vp, i32, ip, ip, ip, dp, dp, ip, dp]),
        "sbh_eigen_cycles": (i32, [vp, i32, i32]), "sbh_eigen_run": (i32, [vp]),
        "sbh_eigen_stats": (i32, [vp, C.POINTER(C.c_longlong), C.POINTER(C.c_longlong), C.POINTER(C.c_longlong), dp]),
        "sbh_eigen_host_bytes": (i32, [vp, i32, C.POINTER(C.c_longlong), C.POINTER(C.c_longlong)]),
        "sbh_model_graph": (i32, [vp, ip, ip]), "sbh_model_xs": (i32, [vp, dp, dp]),
        "sbh_geom_create": (vp, [C.c_char_p, i32, i32]), "sbh_geom_info": (i32, [vp] + [ip] * 8),
        "sbh_geom_uni_fill": (i32, [vp, i32, ip, i32]), "sbh_geom_active_mats": (i32, [vp, ip, i32]),
    }
    for name, (res, args) in sig.items():
        f = getattr(L, name)
        f.restype = res
        f.argtypes = args
    _LIB = L
    return L
