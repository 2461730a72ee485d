"""ctypes binding of libwiflow_b200.so (include/wiflow_b200.h).  No fallback: if the library is missing or the
device is not sm_100a every call raises."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libwiflow_b200.so')

FLAG_TRAIN = 1
FLAG_SAVE = 2
FLAG_PROFILE = 4
(BLOCK_MODEL, BLOCK_TCN, BLOCK_CONVBLOCK1, BLOCK_ASYMCONV, BLOCK_AXIAL_W, BLOCK_AXIAL_H, BLOCK_DUAL_AXIAL,
 BLOCK_INNER_TCN) = range(8)
LOSS_TYPES = {'smooth_l1': 0, 'mse': 1, 'l1': 2}


class BlockDesc(ctypes.Structure):
    _fields_ = [('block', ctypes.c_int), ('cin', ctypes.c_int), ('cout', ctypes.c_int), ('width', ctypes.c_int),
                ('dilation', ctypes.c_int)]

    def key(self):
        return (self.block, self.cin, self.cout, self.width, self.dilation)


_lib = None


def lib():
    """Load (once) and return the shared library; raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f'{LIB_PATH} not found: build it with `python __graft_entry__.py build` '
                           '(nvcc, sm_100a). There is no CPU or PyTorch fallback for the WiFlow kernels.')
    L = ctypes.CDLL(LIB_PATH)
    vp, ip, ll, f = ctypes.c_void_p, ctypes.c_int, ctypes.c_longlong, ctypes.c_float
    dp = ctypes.POINTER(BlockDesc)
    L.wf_last_error_string.restype = ctypes.c_char_p
    L.wf_version.restype = ip
    for name in ('wf_param_count', 'wf_running_count'):
        getattr(L, name).restype = ll
        getattr(L, name).argtypes = [dp]
    for name in ('wf_bn_count', 'wf_dropout_sites'):
        getattr(L, name).restype = ip
        getattr(L, name).argtypes = [dp]
    L.wf_param_table.restype = ip
    L.wf_param_table.argtypes = [dp, ip, ctypes.c_char_p, ip, ctypes.POINTER(ll), ctypes.POINTER(ll)]
    L.wf_workspace_bytes.restype = ctypes.c_size_t
    L.wf_workspace_bytes.argtypes = [dp, ip, ip]
    L.wf_block_forward.restype = ip
    L.wf_block_forward.argtypes = [dp, vp, vp, vp, vp, vp, vp, vp, ctypes.c_size_t, ip, ip, vp]
    L.wf_block_backward.restype = ip
    L.wf_block_backward.argtypes = [dp, vp, vp, vp, vp, vp, vp, vp, ctypes.c_size_t, ip, ip, vp]
    L.wf_pose_loss.restype = ip
    L.wf_pose_loss.argtypes = [vp, vp, ip, ip, f, f, vp, vp, vp, vp, vp]
    L.wf_pose_metrics.restype = ip
    L.wf_pose_metrics.argtypes = [vp, vp, ip, ctypes.POINTER(f), ip, ip, vp, vp, vp]
    L.wf_clip_adamw.restype = ip
    L.wf_clip_adamw.argtypes = [vp, vp, vp, vp, ll, vp, f, f, f, f, f, f, f, vp]
    L.wf_window_load.restype = ip
    L.wf_window_load.argtypes = [vp, ll, vp, vp, ip, ip, ip, ip, vp, vp, vp]
    L.wf_noise_scale.restype = ip
    L.wf_noise_scale.argtypes = [vp, vp, vp, ll, f, f, vp, ll, vp]
    L.wf_dropout_masks.restype = ip
    L.wf_dropout_masks.argtypes = [vp, vp, vp, ip, ctypes.c_ulonglong, vp, vp]
    L.wf_keypoint_batch.restype = ip
    L.wf_keypoint_batch.argtypes = [vp, ll, vp, vp, ip, ip, ip, vp]
    L.wf_keypoint_sequences.restype = ip
    L.wf_keypoint_sequences.argtypes = [vp, vp, ip, ip, vp]
    L.wf_debug_tensor.restype = ip
    L.wf_debug_tensor.argtypes = [dp, ip, ip, ip, ctypes.c_char_p, ip, ctypes.POINTER(ll), ctypes.POINTER(ip), ctypes.POINTER(ip)]
    L.wf_launch_count.restype = ll
    L.wf_profile_count.restype = ip
    L.wf_profile_read.restype = ip
    L.wf_profile_read.argtypes = [ip, ctypes.c_char_p, ip, ctypes.POINTER(f), ctypes.POINTER(ctypes.c_double)]
    L.wf_profile_reset.restype = None
    L.wf_profile_bytes.restype = ip
    L.wf_profile_bytes.argtypes = [ip, ctypes.POINTER(ctypes.c_double)]
    _lib = L
    return L


def check(rc, what):
    if rc != 0:
        msg = lib().wf_last_error_string().decode('utf-8', 'replace')
        raise RuntimeError(f'{what} failed (code {rc}): {msg}')


def param_table(desc):
    """[(name, offset, numel)] of the block's parameters in state_dict order."""
    L = lib()
    out, i = [], 0
    name = ctypes.create_string_buffer(256)
    off, num = ctypes.c_longlong(), ctypes.c_longlong()
    while L.wf_param_table(ctypes.byref(desc), i, name, 256, ctypes.byref(off), ctypes.byref(num)) == 0:
        out.append((name.value.decode(), off.value, num.value))
        i += 1
    return out


def debug_tensors(desc, B, flags):
    """{name: (byte_offset, C, P)} of the named workspace tensors (tests only)."""
    L = lib()
    out, i = {}, 0
    name = ctypes.create_string_buffer(256)
    off, C, P = ctypes.c_longlong(), ctypes.c_int(), ctypes.c_int()
    while L.wf_debug_tensor(ctypes.byref(desc), B, flags, i, name, 256, ctypes.byref(off), ctypes.byref(C), ctypes.byref(P)) == 0:
        out[name.value.decode()] = (off.value, C.value, P.value)
        i += 1
    return out


def profile_records(with_bytes=False):
    """[(name, ms, flops[, algorithmic HBM bytes])] of the launches recorded with FLAG_PROFILE on this thread; clears the records."""
    L = lib()
    out = []
    name = ctypes.create_string_buffer(256)
    ms = ctypes.c_float()
    fl = ctypes.c_double()
    by = ctypes.c_double()
    for i in range(L.wf_profile_count()):
        check(L.wf_profile_read(i, name, 256, ctypes.byref(ms), ctypes.byref(fl)), 'wf_profile_read')
        if with_bytes:
            check(L.wf_profile_bytes(i, ctypes.byref(by)), 'wf_profile_bytes')
            out.append((name.value.decode(), ms.value, fl.value, by.value))
        else:
            out.append((name.value.decode(), ms.value, fl.value))
    L.wf_profile_reset()
    return out
