"""ctypes binding of include/webrtc_ns_b200.h (argument types only; no logic)."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# NSB200_LIB: alternative build of the same library (kernel tuning experiments)
LIB_PATH = os.environ.get("NSB200_LIB") or os.path.join(HERE, "libwebrtc_ns_b200.so")

# every symbol include/webrtc_ns_b200.h declares: (restype, argtypes)
_H = C.c_void_p
_HP = C.POINTER(C.c_void_p)
SYMBOLS = {
    "WebRtcNs_Create": (C.c_int, [_HP]),
    "WebRtcNs_Free": (C.c_int, [_H]),
    "WebRtcNs_Init": (C.c_int, [_H, C.c_uint32]),
    "WebRtcNs_set_policy": (C.c_int, [_H, C.c_int]),
    "WebRtcNs_Analyze": (None, [_H, C.c_void_p]),
    "WebRtcNs_Process": (None, [_H, C.POINTER(C.c_void_p), C.c_int, C.POINTER(C.c_void_p)]),
    "WebRtcNs_prior_speech_probability": (C.c_float, [_H]),
    "WebRtcNsx_Create": (C.c_int, [_HP]),
    "WebRtcNsx_Free": (C.c_int, [_H]),
    "WebRtcNsx_Init": (C.c_int, [_H, C.c_uint32]),
    "WebRtcNsx_set_policy": (C.c_int, [_H, C.c_int]),
    "WebRtcNsx_Process": (None, [_H, C.POINTER(C.c_void_p), C.c_int, C.POINTER(C.c_void_p)]),
    "WebRtcNs_ProcessBatch": (C.c_int, [_HP, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_int]),
    "WebRtcNsx_ProcessBatch": (C.c_int, [_HP, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_int]),
    "WebRtcNs_ProcessBatchAsync": (C.c_int, [_HP, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_int, C.POINTER(C.c_uint64)]),
    "WebRtcNsx_ProcessBatchAsync": (C.c_int, [_HP, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_int, C.POINTER(C.c_uint64)]),
    "WebRtcNsB200_WaitBatch": (C.c_int, [C.c_uint64]),
    "WebRtcNs_ProcessBatchDevice": (C.c_int, [_HP, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]),
    "WebRtcNsx_ProcessBatchDevice": (C.c_int, [_HP, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]),
    "WebRtcNs_ProcessBatchBandsF32": (C.c_int, [_HP, C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_int]),
    "WebRtcNs_AnalyzeProcessBatch": (C.c_int, [_HP, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_int]),
    "WebRtcNs_AnalyzeProcessBatchDevice": (C.c_int, [_HP, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]),
    "WebRtcNs_AnalyzeProcessBatchBandsF32": (C.c_int, [_HP, C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_int]),
    "WebRtcNs_ProcessInterleavedI16": (C.c_int, [_HP, C.c_int, C.c_void_p, C.c_int]),
    "WebRtcNs_ProcessInterleavedF32": (C.c_int, [_HP, C.c_int, C.c_void_p, C.c_int]),
    "WebRtcNs_InitBatch": (C.c_int, [_HP, C.c_int, C.c_uint32, C.c_int]),
    "WebRtcNsx_InitBatch": (C.c_int, [_HP, C.c_int, C.c_uint32, C.c_int]),
    "WebRtcNsB200_StateSize": (C.c_size_t, [_H]),
    "WebRtcNsB200_ExportState": (C.c_int, [_H, C.c_void_p, C.c_size_t]),
    "WebRtcNsB200_ImportState": (C.c_int, [_H, C.c_void_p, C.c_size_t]),
    "WebRtcNsB200_MigrateHandle": (C.c_int, [_H, C.c_int]),
    "WebRtcNsB200_HandleDevice": (C.c_int, [_H]),
    "WebRtcNsB200_SetCreateDevice": (C.c_int, [C.c_int]),
    "WebRtcNsB200_DeviceCount": (C.c_int, []),
    "WebRtcNsB200_Synchronize": (C.c_int, []),
    "WebRtcNsB200_LastError": (C.c_char_p, []),
    "WebRtcNsB200_KernelLaunches": (C.c_uint64, []),
    "WebRtcNsB200_SelfTest": (C.c_int, [C.c_uint64]),
    "WebRtcNsB200_SelfTestStats": (C.c_int, [C.c_uint64, C.POINTER(C.c_uint64)]),
    "WebRtcNsB200_SynthPcmDevice": (C.c_int, [C.c_void_p, C.c_size_t, C.c_int, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p]),
    "WebRtcNsB200_SynthPcmHost": (None, [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32]),
    "WebRtcNsB200_ChecksumDevice": (C.c_int, [C.c_void_p, C.c_size_t, C.c_int, C.c_uint32, C.c_void_p, C.c_void_p]),
    "WebRtcNsB200_ChecksumAccumulateDevice": (C.c_int, [C.c_void_p, C.c_size_t, C.c_int, C.c_uint32, C.c_void_p, C.c_void_p]),
}

_lib = None


def load_library():
    """Loads the CUDA library; raises (never falls back) when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "libwebrtc_ns_b200.so is not built (run `python -m audiosignalprocess_b200.build`); "
                "this package has no CPU implementation")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib
