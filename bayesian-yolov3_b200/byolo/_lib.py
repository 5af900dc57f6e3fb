"""ctypes binding of libbyolo.so (include/byolo.h).  There is no fallback: a missing library is an ImportError-grade
failure for every entry point of the package."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, 'libbyolo.so')

STANDARD, ALEATORIC, EPISTEMIC = 0, 1, 2
VARIANT_ID = {'standard': STANDARD, 'aleatoric': ALEATORIC, 'epistemic': EPISTEMIC}
PREC_FP32, PREC_FP16_SIMT, PREC_FP16, PREC_FP16X3 = 0, 1, 2, 3
PRECISION_ID = {'fp32': PREC_FP32, 'fp16-simt': PREC_FP16_SIMT, 'fp16': PREC_FP16, 'fp16x3': PREC_FP16X3}


class Config(C.Structure):
    _fields_ = [('variant', C.c_int32), ('height', C.c_int32), ('width', C.c_int32), ('cls_cnt', C.c_int32),
                ('max_batch', C.c_int32), ('T', C.c_int32), ('precision', C.c_int32),
                ('standard_test_dropout', C.c_int32), ('drop_prob', C.c_float),
                ('prior_h', C.c_float * 9), ('prior_w', C.c_float * 9)]


# name -> (restype, argtypes); must list every symbol include/byolo.h declares (tests/test_host_logic.py checks)
_P, _I, _U64, _F, _SZ = C.c_void_p, C.c_int32, C.c_uint64, C.c_float, C.c_size_t
SIGNATURES = {
    'byolo_version': (C.c_int, []),
    'byolo_last_error': (C.c_char_p, []),
    'byolo_create': (C.c_int, [C.POINTER(Config), C.POINTER(_P)]),
    'byolo_destroy': (C.c_int, [_P]),
    'byolo_load_weights': (C.c_int, [_P, _P, _SZ]),
    'byolo_output_shape': (C.c_int, [_P, C.POINTER(_I), C.POINTER(_I), C.POINTER(_I), C.POINTER(_I)]),
    'byolo_forward': (C.c_int, [_P, _P, _I, _U64, _I, _P, _P]),
    'byolo_nms': (C.c_int, [_P, _I, _I, _I, _I, _F, _I, _P, _P, _P, _P]),
    'byolo_nms_ex': (C.c_int, [_P, _I, _I, _I, _I, _F, _I, _P, _P, _P, _I, _I, _I, _P]),
    'byolo_class_filter': (C.c_int, [_P, _I, _I, _I, _I, _I, _I, _I, _P, _P]),
    'byolo_detect_packed': (C.c_int, [_P, _P, _I, _U64, _I, _F, _I, _P, _P, _P]),
    'byolo_detect': (C.c_int, [_P, _P, _I, _U64, _I, _F, _I, _P, _P, _P, _P, _P]),
    'byolo_detect_host': (C.c_int, [_P, _P, _I, _U64, _I, _F, _I, _P, _P, _P]),
    'byolo_submit_host': (C.c_int, [_P, _P, _I, _U64, _I, _F, _I, _P, _P, _I, _P]),
    'byolo_wait_host': (C.c_int, [_P, _I]),
    'byolo_decode': (C.c_int, [_P, _P, _P, _P, _I, _P, _P]),
    'byolo_conv_layer': (C.c_int, [_I, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _P, _P, _P, _P, _I, _I, _I, _U64, _I, _F,
                                   _I, _I, _P, _P]),
    'byolo_get_activation': (C.c_int, [_P, _I, _P, _SZ, C.POINTER(_I * 4), _P]),
    'byolo_profile': (C.c_int, [_P, _I]),
    'byolo_profile_read': (C.c_int, [_P, _P, _P, _P, _P, _P, _I]),
    'byolo_profile_read_coarse': (C.c_int, [_P, _P, _P, _P, _I]),
    'byolo_launch_count': (C.c_int, [_P, _I]),
    'byolo_flops_per_image': (C.c_double, [_P]),
    'byolo_flops_per_image_executed': (C.c_double, [_P]),
}

_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError('libbyolo.so is missing (%s): build it with `python -m byolo.build` / '
                               '__graft_entry__.build(); there is no CPU or PyTorch fallback' % LIB_PATH)
        h = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(h, name)
            fn.restype, fn.argtypes = res, args
        _LIB = h
    return _LIB


class ByoloError(RuntimeError):
    pass


def check(rc):
    if rc < 0:
        raise ByoloError('libbyolo error %d: %s' % (rc, lib().byolo_last_error().decode()))
    return rc
