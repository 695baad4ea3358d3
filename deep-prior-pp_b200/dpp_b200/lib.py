"""ctypes loader for libdpp_b200.so (C ABI declared in include/dpp_b200.h)."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.normpath(os.path.join(_HERE, '..', 'csrc', 'libdpp_b200.so'))
if os.environ.get('DPP_LIB'):      # tools/conv_probe.py: the -DDPP_PROFILE debug build
    _LIB_PATH = os.path.abspath(os.environ['DPP_LIB'])


class DppError(RuntimeError):
    pass


def library_path():
    return _LIB_PATH


class BnRef(C.Structure):
    _fields_ = [('sums', C.c_void_p), ('mean', C.c_void_p), ('inv_std', C.c_void_p),
                ('gamma', C.c_void_p), ('beta', C.c_void_p), ('count', C.c_double),
                ('eps', C.c_float), ('relu', C.c_int)]


class AugRec(C.Structure):
    _fields_ = [('src_index', C.c_int32), ('mode', C.c_int32), ('half_old', C.c_float),
                ('comz_old', C.c_float), ('zstart', C.c_float), ('zend', C.c_float),
                ('bg', C.c_float), ('lo', C.c_float), ('comz_new', C.c_float),
                ('half_new', C.c_float), ('m', C.c_double * 9)]


class CropRec(C.Structure):
    _fields_ = [('src_index', C.c_int32), ('xstart', C.c_int32), ('ystart', C.c_int32), ('wb', C.c_int32),
                ('hb', C.c_int32), ('rw', C.c_int32), ('rh', C.c_int32), ('px', C.c_int32), ('py', C.c_int32),
                ('flags', C.c_int32), ('zstart', C.c_float), ('zend', C.c_float), ('fill', C.c_float),
                ('hi', C.c_float), ('lo', C.c_float), ('comz', C.c_float), ('half', C.c_float),
                ('reserved', C.c_float), ('ifx', C.c_double), ('ify', C.c_double)]


class ConvDesc(C.Structure):
    _fields_ = [('N', C.c_int), ('H', C.c_int), ('W', C.c_int), ('Cin', C.c_int),
                ('Cout', C.c_int), ('k', C.c_int), ('stride', C.c_int), ('pad', C.c_int),
                ('Ho', C.c_int), ('Wo', C.c_int), ('precision', C.c_int),
                ('wpack_fwd', C.c_void_p), ('wpack_dgrad', C.c_void_p)]


class PackItem(C.Structure):
    _fields_ = [('w', C.c_void_p), ('img_fwd', C.c_void_p), ('img_dgrad', C.c_void_p), ('Cin', C.c_int),
                ('Cout', C.c_int), ('k', C.c_int), ('bn_fwd', C.c_int), ('bn_dgrad', C.c_int), ('passes', C.c_int)]


class WgradLayer(C.Structure):
    _fields_ = [('d', ConvDesc), ('x', C.c_void_p), ('in_bn', BnRef), ('has_in_bn', C.c_int), ('dy', C.c_void_p),
                ('dw', C.c_void_p), ('db', C.c_void_p)]


class BnEmaItem(C.Structure):
    _fields_ = [('sums', C.c_void_p), ('mean', C.c_void_p), ('inv_std', C.c_void_p),
                ('count', C.c_double), ('C', C.c_int), ('eps', C.c_float)]


# numpy dtype mirroring dpp_aug_rec (for vectorised host-side record preparation)
AUG_REC_DTYPE = [('src_index', '<i4'), ('mode', '<i4'), ('half_old', '<f4'), ('comz_old', '<f4'),
                 ('zstart', '<f4'), ('zend', '<f4'), ('bg', '<f4'), ('lo', '<f4'),
                 ('comz_new', '<f4'), ('half_new', '<f4'), ('m', '<f8', (9,))]

# numpy dtype mirroring dpp_crop_rec
CROP_REC_DTYPE = [('src_index', '<i4'), ('xstart', '<i4'), ('ystart', '<i4'), ('wb', '<i4'), ('hb', '<i4'),
                  ('rw', '<i4'), ('rh', '<i4'), ('px', '<i4'), ('py', '<i4'), ('flags', '<i4'), ('zstart', '<f4'),
                  ('zend', '<f4'), ('fill', '<f4'), ('hi', '<f4'), ('lo', '<f4'), ('comz', '<f4'), ('half', '<f4'),
                  ('reserved', '<f4'), ('ifx', '<f8'), ('ify', '<f8')]
CROP_NORMALISE, CROP_CLAMP, CROP_MIRROR = 1, 2, 4

DPP_ENOTSUP = -3

P = C.c_void_p
_SIGS = {
    'dpp_abi_version': (C.c_int, []),
    'dpp_last_error': (C.c_char_p, []),
    'dpp_device_info': (C.c_int, [C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_char_p, C.c_int]),
    'dpp_nchw_to_nhwc': (C.c_int, [P, P, C.c_int, C.c_int, C.c_int, C.c_int, P]),
    'dpp_nhwc_to_nchw': (C.c_int, [P, P, C.c_int, C.c_int, C.c_int, C.c_int, P]),
    'dpp_augment_fwd': (C.c_int, [P, P, P, C.c_int, C.c_int, C.c_int, P]),
    'dpp_recrop_fwd': (C.c_int, [P, P, P, P, P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, P]),
    'dpp_sample_poses': (C.c_int, [P] * 8 + [C.c_double] * 4 + [C.c_int, P, P, P, C.c_int, C.c_int, P]),
    'dpp_joint_errors': (C.c_int, [P, P, P, P, P, C.c_int, C.c_int, P]),
    'dpp_convpool_fwd': (C.c_int, [P, P, P, P, P, P] + [C.c_int] * 9 + [P]),
    'dpp_convpool_bwd': (C.c_int, [P, P, P, P, P, P, P, P] + [C.c_int] * 9 + [P]),
    'dpp_conv_pack_size': (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    'dpp_conv_pack_all': (C.c_int, [P, C.c_int, P]),
    'dpp_conv2d_fwd': (C.c_int, [C.POINTER(ConvDesc), P, C.POINTER(BnRef), P, P, P, P, P, P]),
    'dpp_conv2d_dgrad': (C.c_int, [C.POINTER(ConvDesc), P, P, P, C.c_int, C.POINTER(BnRef), P, P, P]),
    'dpp_conv2d_dgrad_bn_bwd': (C.c_int, [C.POINTER(ConvDesc), P, P, P, C.c_int, C.POINTER(BnRef), P, P, P, P, P, P,
                                          C.c_float, P, P]),
    'dpp_conv2d_wgrad': (C.c_int, [C.POINTER(ConvDesc), P, C.POINTER(BnRef), P, P, P, P]),
    'dpp_wgrad_group_create': (C.c_int, [P, C.c_int, C.POINTER(C.c_void_p)]),
    'dpp_wgrad_group_run': (C.c_int, [P, P]),
    'dpp_wgrad_group_launches': (C.c_int, [P]),
    'dpp_wgrad_group_destroy': (C.c_int, [P]),
    'dpp_bn_bwd_apply': (C.c_int, [P, P, C.POINTER(BnRef), P, P, P, P, P, P, C.c_int64, C.c_int, C.c_float, P]),
    'dpp_peer_alloc': (C.c_int, [C.c_int64, C.POINTER(C.c_void_p), C.c_char_p]),
    'dpp_peer_open': (C.c_int, [C.c_char_p, C.POINTER(C.c_void_p)]),
    'dpp_peer_close': (C.c_int, [P]),
    'dpp_peer_free': (C.c_int, [P]),
    'dpp_stats_exchange': (C.c_int, [P, C.c_int, P, C.c_int64, C.c_int, C.c_int, P, P, P]),
    'dpp_bn_apply': (C.c_int, [P, C.POINTER(BnRef), P, C.c_int64, C.c_int, P]),
    'dpp_bn_relu_bwd_reduce': (C.c_int, [P, P, C.POINTER(BnRef), P, P, C.c_int64, C.c_int, P]),
    'dpp_bn_ema_update': (C.c_int, [P, C.c_int, C.c_float, P]),
    'dpp_fc_fwd': (C.c_int, [P, P, P, P, C.c_int, C.c_int, C.c_int, C.c_int, P, C.c_float, C.c_int, P]),
    'dpp_fc_bwd': (C.c_int, [P, P, P, P, P, P, P, P, C.c_int, C.c_int, C.c_int, C.c_int, P, C.c_float,
                             C.c_int, P]),
    'dpp_fc_bwd_ex': (C.c_int, [P, P, P, P, P, P, P, P, C.c_int, C.c_int, C.c_int, C.c_int, P, C.c_float,
                                C.c_int, C.c_int, P]),
    'dpp_fc_workspace_init': (C.c_int, []),
    'dpp_loss_sqerr': (C.c_int, [P, P, P, P, C.c_int, C.c_int, P]),
    'dpp_adam_step': (C.c_int, [P, P, P, P, P, C.c_int64, P]),
    'dpp_adam_tick': (C.c_int, [P, P]),
    'dpp_set_pdl': (C.c_int, [C.c_int]),
    'dpp_wgrad_workspace_init': (C.c_int, []),
    'dpp_copy2d': (C.c_int, [P, C.c_int64, P, C.c_int64, C.c_int64, C.c_int64, P]),
    'dpp_fill_f32': (C.c_int, [P, C.c_float, C.c_int64, P]),
    'dpp_fill_f64': (C.c_int, [P, C.c_double, C.c_int64, P]),
}

EXPORTED_SYMBOLS = sorted(_SIGS)


class _Lib(object):
    """Lazy handle: ``lib.dpp_xxx(...)`` raises DppError on a non-zero return code."""

    def __init__(self):
        self._dll = None

    def load(self):
        if self._dll is None:
            if not os.path.exists(_LIB_PATH):
                raise DppError("libdpp_b200.so is not built (%s). Run `python -c 'import __graft_entry__ as g; "
                               "g.build()'` or `make -C deep-prior-pp_b200/csrc`; there is no CPU fallback." % _LIB_PATH)
            dll = C.CDLL(_LIB_PATH)
            for name, (res, args) in _SIGS.items():
                fn = getattr(dll, name)
                fn.restype = res
                fn.argtypes = args
            self._dll = dll
        return self._dll

    def raw(self, name):
        return getattr(self.load(), name)

    def __getattr__(self, name):
        if name.startswith('_'):
            raise AttributeError(name)
        fn = getattr(self.load(), name)
        if name in ('dpp_last_error', 'dpp_abi_version', 'dpp_wgrad_group_launches'):
            return fn

        def call(*a):
            rc = fn(*a)
            if rc != 0:
                raise DppError("%s failed (%d): %s" % (name, rc, self._dll.dpp_last_error().decode()))
            return rc
        call.__name__ = name
        return call


lib = _Lib()
