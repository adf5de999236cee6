"""ctypes binding of libdeepxi_b200.so (C ABI: include/deepxi_b200.h).

PyTorch is used only to own device memory and streams: every function here takes torch CUDA tensors,
checks dtype / contiguity, and passes `data_ptr()` and the current stream handle across the C ABI.
There is no CPU fallback: if the library is missing, or the tensors are not on a CUDA device, the
calls raise.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('DXI_LIB') or os.path.join(_HERE, 'libdeepxi_b200.so')      # DXI_LIB: the tuning build (deepxi_b200/build.py)

GTYPES = {'mmse-lsa': 0, 'mmse-stsa': 1, 'wf': 2, 'srwf': 3, 'cwf': 4, 'irm': 5, 'ibm': 6, 'deepmmse': 7}
NET_KINDS = {'ResNetV2': 0, 'MHANetV3': 1, 'ResNet': 2, 'ResNetV3': 3}
PRECISIONS = {'f32': 0, 'f16x3': 1, 'f16': 2}
PADDINGS = {'causal': 0, 'same': 1}
MASK_MODES = {'none': 0, 'causal+pad': 1}

# every symbol include/deepxi_b200.h declares
SYMBOLS = ['dxi_last_error', 'dxi_host_alloc', 'dxi_host_free', 'dxi_version', 'dxi_device_check', 'dxi_stft', 'dxi_istft', 'dxi_map_gain', 'dxi_gfunc',
           'dxi_cdf_map', 'dxi_deepmmse', 'dxi_enhance', 'dxi_net_create', 'dxi_net_load', 'dxi_net_finalize',
           'dxi_net_workspace_bytes', 'dxi_net_forward', 'dxi_net_destroy', 'dxi_launch_count',
           'dxi_launch_count_reset', 'dxi_selftest_umma', 'dxi_selftest_umma_pair', 'dxi_profile_enable', 'dxi_profile_read',
           'dxi_mix_workspace_bytes', 'dxi_mix', 'dxi_xi_map', 'dxi_xi_db_moments', 'dxi_subband_ibm']


class DxiError(RuntimeError):
    pass


class NetCfg(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in ('n_feat', 'n_outp', 'd_model', 'n_blocks', 'd_f', 'k', 'max_d_rate',
                                              'padding', 'n_heads', 'max_len', 'mask_mode', 'precision')]


_lib = None


def load():
    """Loads the shared library (once).  Raises if it has not been built: there is no other code path."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise DxiError('%s not found: build it with `python -m deepxi_b200.build` (no CPU fallback exists)' % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    vp, i32, i64 = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64
    lib.dxi_last_error.restype = ctypes.c_char_p
    lib.dxi_last_error.argtypes = []
    lib.dxi_version.restype = i32
    lib.dxi_host_alloc.argtypes = [ctypes.POINTER(vp), ctypes.c_size_t, i32]
    lib.dxi_host_alloc.restype = i32
    lib.dxi_host_free.argtypes = [vp]
    lib.dxi_host_free.restype = i32
    lib.dxi_device_check.restype = i32
    lib.dxi_stft.argtypes = [vp, i32, vp, i32, i64, i32, vp, vp, vp]
    lib.dxi_istft.argtypes = [vp, vp, vp, vp, i32, i32, vp, vp, i64, vp]
    lib.dxi_map_gain.argtypes = [vp, vp, vp, i64, i32, i32, vp, vp, vp, vp]
    lib.dxi_gfunc.argtypes = [vp, vp, i64, i32, vp, vp]
    lib.dxi_cdf_map.argtypes = [vp, vp, vp, i64, i32, vp, vp]
    lib.dxi_deepmmse.argtypes = [vp, vp, vp, vp, i64, i32, vp, vp]
    lib.dxi_deepmmse.restype = i32
    lib.dxi_enhance.argtypes = [vp, vp, vp, vp, vp, i32, vp, i32, i32, vp, vp, i64, vp]
    lib.dxi_net_create.argtypes = [ctypes.POINTER(vp), i32, ctypes.POINTER(NetCfg)]
    lib.dxi_net_load.argtypes = [vp, ctypes.c_char_p, vp, ctypes.POINTER(i64), i32]
    lib.dxi_net_finalize.argtypes = [vp, vp]
    lib.dxi_net_workspace_bytes.argtypes = [vp, i32, i32]
    lib.dxi_net_workspace_bytes.restype = i64
    lib.dxi_net_forward.argtypes = [vp, vp, i32, i32, vp, vp, ctypes.c_size_t, vp]
    lib.dxi_net_destroy.argtypes = [vp]
    lib.dxi_launch_count.restype = i64
    lib.dxi_launch_count.argtypes = []
    lib.dxi_launch_count_reset.restype = None
    lib.dxi_selftest_umma.argtypes = [vp, vp, i32, i32, i32, vp, vp]
    lib.dxi_selftest_umma_pair.argtypes = [vp, vp, i32, vp, vp]
    lib.dxi_profile_enable.argtypes = [i32]
    lib.dxi_profile_enable.restype = None
    lib.dxi_profile_read.argtypes = [ctypes.c_char_p, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(i64)]
    lib.dxi_profile_read.restype = i32
    if hasattr(lib, 'dxi_debug_tcn_clocks'):      # tuning build only (include/deepxi_b200_debug.h)
        lib.dxi_debug_tcn_clocks.argtypes = [vp, i32]
        lib.dxi_debug_tcn_clocks.restype = None
        lib.dxi_debug_tmem_bw.argtypes = [i32, i32, i32, vp, vp]
        lib.dxi_debug_tmem_bw.restype = i32
        lib.dxi_debug_tcn_stop_after.argtypes = [i32]
        lib.dxi_debug_tcn_stop_after.restype = None
        lib.dxi_debug_chain_clocks.argtypes = [vp, i32]
        lib.dxi_debug_chain_clocks.restype = i32
    lib.dxi_mix_workspace_bytes.argtypes = [i32]
    lib.dxi_mix_workspace_bytes.restype = i64
    lib.dxi_mix.argtypes = [vp, vp, vp, vp, vp, vp, i32, i64, i64, vp, vp, vp, i64, vp, vp]
    lib.dxi_xi_map.argtypes = [vp, vp, vp, vp, i64, i32, vp, vp, vp]
    lib.dxi_xi_db_moments.argtypes = [vp, vp, vp, i32, i32, i32, vp, vp]
    lib.dxi_subband_ibm.argtypes = [vp, vp, i64, i32, i32, vp, vp, vp]
    for name in ('dxi_mix', 'dxi_xi_map', 'dxi_xi_db_moments', 'dxi_subband_ibm'):
        getattr(lib, name).restype = i32
    for name in ('dxi_stft', 'dxi_istft', 'dxi_map_gain', 'dxi_gfunc', 'dxi_cdf_map', 'dxi_deepmmse', 'dxi_enhance', 'dxi_net_create',
                 'dxi_net_load', 'dxi_net_finalize', 'dxi_net_forward', 'dxi_net_destroy', 'dxi_selftest_umma'):
        getattr(lib, name).restype = i32
    _lib = lib
    return lib


def check(rc, value_error=False):
    if rc != 0:
        msg = load().dxi_last_error().decode()
        if rc == -1 and value_error:
            raise ValueError(msg)
        raise DxiError('libdeepxi_b200 error %d: %s' % (rc, msg))


def stream_ptr(device=None):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def ptr(t, dtype=None, allow_none=False):
    """Device pointer of a contiguous CUDA tensor (None -> NULL)."""
    if t is None:
        if allow_none:
            return ctypes.c_void_p(0)
        raise ValueError('tensor required')
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise DxiError('expected a CUDA tensor: the deepxi_b200 kernels have no CPU path')
    if dtype is not None and t.dtype != dtype:
        raise ValueError('expected dtype %s, got %s' % (dtype, t.dtype))
    if not t.is_contiguous():
        raise ValueError('tensor must be contiguous')
    return ctypes.c_void_p(t.data_ptr())


def gtype_code(gtype):
    """gain.py:190: unknown types raise ValueError('Invalid gain function type.')."""
    if gtype == 'dgwf':
        raise NotImplementedError("'dgwf' needs the constructive-deconstructive mask of the STDCT target (out of scope)")
    if gtype not in GTYPES:
        raise ValueError('Invalid gain function type.')
    return GTYPES[gtype]


def profile_enable(on=True):
    load().dxi_profile_enable(int(bool(on)))


def profile_read(key):
    """(total milliseconds, launches) accumulated for a named kernel group since the last read."""
    ms, n = ctypes.c_double(0.0), ctypes.c_int64(0)
    check(load().dxi_profile_read(key.encode(), ctypes.byref(ms), ctypes.byref(n)))
    return ms.value, n.value


def launch_count():
    return int(load().dxi_launch_count())


def launch_count_reset():
    load().dxi_launch_count_reset()
