"""ctypes loader for libatomorph_b200.so (the C-ABI of include/amx.h and include/amx_morph.h).

There is NO fallback: if the shared object is missing or a CUDA device is absent, calls raise.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("AMX_LIB") or os.path.join(_HERE, "libatomorph_b200.so")     # AMX_LIB: a tuning build (build.py AMX_BUILD_TAG)

AMX_OK = 0
STATUS = {0: "AMX_OK", 1: "AMX_ERR_CUDA", 2: "AMX_ERR_ARG", 3: "AMX_ERR_STATE", 4: "AMX_ERR_NOMEM", 5: "AMX_ERR_BUSY"}

_lib = None


class AmxError(RuntimeError):
    pass


def lib():
    """Load the library (once).  Raises if it has not been built -- never falls back to CPU code."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise AmxError("libatomorph_b200.so is not built: run `python -m atomorph_b200.build` "
                       "(there is no CPU fallback for the morph pipeline)")
    L = C.CDLL(LIB_PATH)
    vp, u64, u32, i32, f64 = C.c_void_p, C.c_uint64, C.c_uint32, C.c_int32, C.c_double
    P = C.POINTER

    def sig(name, res, *args):
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = list(args)

    sig("amx_create", i32, P(vp), i32)
    sig("amx_destroy", None, vp)
    sig("amx_last_error", C.c_char_p, vp)
    sig("amx_version", C.c_char_p)
    sig("amx_set_stream", i32, vp, vp)
    sig("amx_get_stream", vp, vp)
    sig("amx_device_sync", i32, vp)
    sig("amx_set_param", i32, vp, i32, f64)
    sig("amx_get_param", f64, vp, i32)
    sig("amx_reset", i32, vp)
    sig("amx_set_canvas", i32, vp, u32, u32, u32, u32, vp)
    sig("amx_set_frame_count", i32, vp, u32, vp)
    sig("amx_upload_frame", i32, vp, u32, vp, vp, vp)
    sig("amx_upload_frame_device", i32, vp, u32, vp, vp, vp)
    sig("amx_download_fetch", i32, vp, u32, vp)
    sig("amx_download_stored", i32, vp, u32, vp)
    sig("amx_step", i32, vp, u64)
    sig("amx_next_state", i32, vp)
    sig("amx_get_state", C.c_uint, vp)
    sig("amx_get_energy", f64, vp)
    sig("amx_blobify", i32, vp)
    sig("amx_blob_count", i32, vp, u32, P(u32))
    sig("amx_export_blobs", i32, vp, u32, vp, vp, vp)
    sig("amx_import_blobs", i32, vp, u32, u32, vp, vp, vp)
    sig("amx_match_init", i32, vp)
    sig("amx_match_rounds", i32, vp, u64)
    sig("amx_match_energy", i32, vp, P(f64))
    sig("amx_init_chains", i32, vp)
    sig("amx_chain_count", i32, vp, P(u32))
    sig("amx_chain_info", i32, vp, u32, vp)
    sig("amx_export_chain", i32, vp, u32, vp)
    sig("amx_import_chains", i32, vp, u32, vp, vp, vp, u32, vp)
    sig("amx_table_device_ptr", i32, vp, u32, P(vp), P(u64))
    sig("amx_swap_rounds", i32, vp, i32, i32, u64, vp)
    sig("amx_swap_stats", i32, vp, vp)
    sig("amx_swap_rounds_sharded", i32, vp, u32, i32, u64, u64, u64)
    sig("amx_pack_owned", i32, vp, u32, u32, u64, u64, vp, P(u64))
    sig("amx_unpack_owned", i32, vp, u32, u32, u64, u32, vp)
    sig("amx_swap_tiled_epoch", i32, vp, u32, u32, u64, u32, u32, u32)
    sig("amx_swap_local_epoch", i32, vp, u32, u32, u64, u32)
    sig("amx_set_swap_locality", i32, vp, u32)
    sig("amx_pack_tiled", i32, vp, u32, u32, u64, u32, u32, vp, P(u64))
    sig("amx_unpack_tiled", i32, vp, u32, u32, u64, vp)
    sig("amx_cost", i32, vp, P(f64))
    sig("amx_comm_unique_id", i32, vp)
    sig("amx_comm_init", i32, vp, vp, u32, u32)
    sig("amx_comm_destroy", i32, vp)
    sig("amx_comm_enable_p2p", i32, vp)
    sig("amx_comm_disable_p2p", i32, vp)
    sig("amx_comm_info", i32, vp, vp)
    sig("amx_comm_check", i32, vp)
    sig("amx_table_broadcast", i32, vp, u32)
    sig("amx_swap_part_step", i32, vp, u32, u32, u64, u32, u32)
    sig("amx_swap_columns_step", i32, vp, i32, u32, u64, u32, u32)
    sig("amx_swap_phase_count", u32, vp)
    sig("amx_column_hash", i32, vp, u32, vp)
    sig("amx_render_prepare", i32, vp)
    sig("amx_render", i32, vp, vp, u32, vp, i32)
    sig("amx_render_blob", i32, vp, u32, f64, u64, vp, vp, P(C.c_int64), P(u64))
    sig("amx_render_stats", i32, vp, vp)
    sig("amx_render_path_frames", i32, vp, vp)
    sig("amx_kernel_times", i32, vp, i32, vp, vp, vp)
    sig("amx_render_tiled_stats", i32, vp, vp)
    sig("amx_render_pixels", i32, vp, f64, vp)
    sig("amx_set_lookahead", i32, vp, i32)
    sig("amx_lookahead_stats", i32, vp, vp)
    sig("amx_background", i32, vp, f64, vp, i32)
    sig("amx_fluid_create", i32, vp, u32, u32, u32)
    sig("amx_fluid_set_particles", i32, vp, u32, vp)
    sig("amx_fluid_get_particles", i32, vp, u32, vp)
    sig("amx_fluid_step", i32, vp, u64, f64, f64)
    sig("amx_fluid_get_nodes", i32, vp, vp)
    sig("amx_launch_count", u64, vp)
    sig("amx_timer_start", i32, vp)
    sig("amx_timer_stop", i32, vp, P(C.c_float))
    _lib = L
    return L


# every symbol include/amx.h declares (checked by tests/test_abi.py without a GPU)
AMX_SYMBOLS = [
    "amx_create", "amx_destroy", "amx_last_error", "amx_version", "amx_set_stream", "amx_get_stream", "amx_device_sync",
    "amx_set_param", "amx_get_param", "amx_reset", "amx_set_canvas", "amx_set_frame_count", "amx_upload_frame",
    "amx_upload_frame_device", "amx_download_fetch", "amx_download_stored", "amx_step", "amx_next_state",
    "amx_get_state", "amx_get_energy", "amx_blobify", "amx_blob_count", "amx_export_blobs", "amx_import_blobs",
    "amx_match_init", "amx_match_rounds", "amx_match_energy", "amx_init_chains", "amx_chain_count",
    "amx_chain_info", "amx_export_chain", "amx_import_chains", "amx_table_device_ptr", "amx_swap_rounds",
    "amx_swap_stats", "amx_swap_rounds_sharded", "amx_pack_owned", "amx_unpack_owned", "amx_swap_tiled_epoch", "amx_swap_local_epoch", "amx_set_swap_locality", "amx_pack_tiled", "amx_unpack_tiled", "amx_cost", "amx_comm_unique_id", "amx_comm_init", "amx_comm_destroy", "amx_comm_enable_p2p", "amx_comm_disable_p2p", "amx_comm_info", "amx_comm_check", "amx_table_broadcast", "amx_swap_part_step", "amx_swap_columns_step", "amx_swap_phase_count", "amx_column_hash", "amx_render_prepare", "amx_render", "amx_render_blob", "amx_render_stats", "amx_render_path_frames", "amx_kernel_times", "amx_render_tiled_stats", "amx_render_pixels", "amx_set_lookahead", "amx_lookahead_stats", "amx_background",
    "amx_fluid_create", "amx_fluid_set_particles", "amx_fluid_get_particles", "amx_fluid_step",
    "amx_fluid_get_nodes", "amx_launch_count", "amx_timer_start", "amx_timer_stop",
]
