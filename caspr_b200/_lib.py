"""ctypes binding of libcaspr_b200.so (the C ABI declared in include/caspr_b200.h).

The product path has NO fallback: if the library is missing, importing this module raises and
every model call fails loudly.  Build it with ``python -m caspr_b200.build`` (or
``__graft_entry__.build()``).
"""
import ctypes
import os
from ctypes import (c_longlong, POINTER, Structure, c_char_p, c_double, c_float, c_int, c_int32, c_size_t,
                    c_void_p)

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'lib', 'libcaspr_b200.so')


class CasprError(RuntimeError):
    def __init__(self, status, where):
        self.status = status
        super().__init__('%s failed: %s (status %d)' % (where, status_string(status), status))


class CnfWeights(Structure):
    _fields_ = [('W', c_void_p * 4), ('b', c_void_p * 4), ('Wgate', c_void_p * 4),
                ('bgate', c_void_p * 4), ('Wbias', c_void_p * 4), ('hidden', c_int),
                ('ctx_dim', c_int)]


class GnFold(Structure):
    _fields_ = [('table', c_void_p), ('rows_per_sample', c_int), ('relu', c_int)]


class GnStats(Structure):
    _fields_ = [('stats', c_void_p), ('rows_per_sample', c_int), ('groups', c_int), ('extrema', c_void_p)]


ALLREDUCE_FN = ctypes.CFUNCTYPE(c_int, c_void_p, c_int, c_void_p, c_void_p)


class CnfSync(Structure):
    _fields_ = [('n_global', ctypes.c_longlong), ('stage', c_void_p), ('allreduce_sum', ALLREDUCE_FN),
                ('user', c_void_p)]


class MbnParams(Structure):
    _fields_ = [('weight', c_void_p), ('bias', c_void_p), ('running_mean', c_void_p),
                ('running_var', c_void_p)]


# name -> (restype, argtypes); mirrors include/caspr_b200.h one to one
_P = c_void_p
SIGNATURES = {
    'caspr_version': (c_int, []),
    'caspr_build_arch': (c_char_p, []),
    'caspr_status_string': (c_char_p, [c_int]),
    'caspr_launch_count': (ctypes.c_ulonglong, []),
    'caspr_launch_count_add': (None, [ctypes.c_ulonglong]),
    'caspr_profile_enable': (None, [c_int]),
    'caspr_profile_read': (c_int, [c_int, POINTER(c_double), POINTER(ctypes.c_longlong)]),
    'caspr_fps': (c_int, [_P, c_int, c_int, c_int, _P, _P, _P]),
    'caspr_ball_query2': (c_int, [_P, _P, c_int, c_int, c_int, c_float, c_int, _P, c_float, c_int, _P, _P]),
    'caspr_group_points': (c_int, [_P, _P, _P, c_int, _P, c_int, c_int, c_int, c_int, c_int, _P, c_int, _P]),
    'caspr_three_nn': (c_int, [_P, _P, c_int, c_int, c_int, _P, _P, _P]),
    'caspr_three_interp_concat': (c_int, [_P, c_int, _P, _P, _P, c_int, c_int, c_int, c_int, c_int, c_int,
                                          _P, c_int, _P]),
    'caspr_linear': (c_int, [_P, c_int, _P, c_int, _P, _P, c_int, c_int, c_int, c_int, c_int, c_int, _P]),
    'caspr_linear_gn_ball': (c_int, [_P, c_int, _P, c_int, _P, _P, _P, c_float, c_int, c_int, c_int, c_int, c_int,
                                     _P, c_int, _P, c_int, _P]),
    'caspr_sa_fused_supported': (c_int, [c_int, c_int, c_int, c_int, c_int]),
    'caspr_sa_fused': (c_int, [_P, _P, _P, c_int, c_int, _P, c_int, c_int, c_int, c_int,
                               _P, _P, _P, _P, c_int, _P, _P, _P, _P, c_int, _P, _P, _P, _P, c_int,
                               c_float, _P, c_int, _P]),
    'caspr_sa_mma_supported': (c_int, [c_int, c_int, c_int, c_int, c_int]),
    'caspr_sa_absmax': (c_int, [_P, _P, c_int, c_int, c_int, c_int, _P, _P]),
    'caspr_sa_mma': (c_int, [_P, _P, _P, c_int, c_int, _P, c_int, c_int, c_int, c_int,
                             _P, _P, _P, _P, c_int, _P, _P, _P, _P, c_int, _P, _P, _P, _P, c_int,
                             c_float, _P, _P, c_int, _P]),
    'caspr_linear_tc_workspace_bytes': (c_size_t, [c_int, c_int, c_int]),
    'caspr_linear_tc_weight_bytes': (c_size_t, [c_int, c_int]),
    'caspr_linear_tc_prepare_weights': (c_int, [_P, c_int, c_int, c_int, _P, c_size_t, _P]),
    'caspr_gn_table': (c_int, [_P, c_int, c_int, c_int, c_int, c_float, _P, _P, _P, _P]),
    'caspr_linear_tc': (c_int, [_P, c_int, _P, c_int, _P, _P, c_int, c_int, c_int, c_int, c_int, c_int, _P,
                                POINTER(GnFold), POINTER(GnStats), c_int, _P, c_size_t, _P]),
    'caspr_groupnorm': (c_int, [_P, c_int, c_int, c_int, c_int, c_int, _P, _P, c_float, c_int, c_int, _P,
                                c_int, _P, c_int, _P]),
    'caspr_groupnorm_project': (c_int, [_P, c_int, c_int, c_int, c_int, c_int, _P, _P, c_float, _P, c_int, _P, _P, _P,
                                        c_int, c_int, _P, c_int, _P]),
    'caspr_gn_max_from_extrema': (c_int, [_P, _P, c_int, c_int, c_int, c_int, _P, _P, c_float, _P, c_int, _P]),
    'caspr_augment_xyz': (c_int, [_P, c_int, _P, _P]),
    'caspr_strip_time': (c_int, [_P, c_int, _P, _P]),
    'caspr_broadcast_rows': (c_int, [_P, c_int, c_int, c_int, c_int, _P, c_int, _P]),
    'caspr_latent_ode_workspace_bytes': (c_size_t, [c_int, c_int, c_int]),
    'caspr_latent_ode_solve': (c_int, [_P, c_int, c_int, c_int, _P, _P, _P, _P, _P, _P, _P, _P,
                                       POINTER(c_double), c_int, c_float, c_float, _P, _P,
                                       POINTER(c_int32), _P, c_size_t, _P]),
    'caspr_cnf_workspace_bytes': (c_size_t, [c_int, c_int, c_int, c_int, c_int]),
    'caspr_cnf_flow': (c_int, [_P, _P, _P, _P, c_int, c_int, POINTER(CnfWeights), POINTER(MbnParams),
                               POINTER(MbnParams), c_float, c_int, c_float, c_float, c_int, _P, _P, _P,
                               POINTER(c_int32), _P, c_size_t, _P]),
    'caspr_cnf_flow_lockstep': (c_int, [_P, _P, _P, _P, c_int, c_int, POINTER(CnfWeights), POINTER(MbnParams),
                                        POINTER(MbnParams), c_float, c_int, c_float, c_float, c_int, _P, _P, _P,
                                        POINTER(c_int32), _P, c_size_t, _P, POINTER(CnfSync)]),
    'caspr_cnf_feval': (c_int, [_P, _P, _P, c_int, c_int, POINTER(CnfWeights), c_float, c_int, _P, _P,
                                _P, c_size_t, _P]),
    'caspr_cnf_param_count': (c_size_t, [c_int, c_int]),
    'caspr_cnf_adjoint_workspace_bytes': (c_size_t, [c_int, c_int, c_int, c_int]),
    'caspr_cnf_adjoint': (c_int, [_P, _P, _P, _P, _P, _P, c_int, c_int, POINTER(CnfWeights), c_float, c_float,
                                  c_float, c_int, _P, _P, _P, _P, _P, _P, POINTER(c_int32), _P, c_size_t, _P]),
    'caspr_latent_ode_param_count': (c_size_t, [c_int, c_int]),
    'caspr_latent_ode_adjoint_workspace_bytes': (c_size_t, [c_int, c_int, c_int]),
    'caspr_latent_ode_adjoint': (c_int, [_P, _P, c_int, c_int, c_int, _P, _P, _P, _P, _P, _P, _P, _P,
                                         POINTER(c_double), c_int, c_float, c_float, _P, _P, _P,
                                         POINTER(c_int32), _P, c_size_t, _P]),
    'caspr_gn_workspace_bytes': (c_size_t, [c_int, c_int, c_int]),
    'caspr_gn_moments': (c_int, [_P, c_int, c_int, c_int, c_int, c_int, c_float, _P, _P, c_size_t, _P]),
    'caspr_gn_apply': (c_int, [_P, c_int, _P, c_int, c_int, c_int, c_int, _P, _P, c_int, _P, c_int, _P]),
    'caspr_rowmax_workspace_bytes': (c_size_t, [c_int, c_int, c_int]),
    'caspr_rowmax': (c_int, [_P, c_int, c_int, c_int, c_int, _P, c_int, _P, _P, c_size_t, _P]),
    'caspr_gn_backward': (c_int, [_P, c_int, _P, c_int, _P, _P, c_int, _P, c_int, c_int, c_int, c_int, _P, _P,
                                  c_int, _P, c_int, _P, _P, _P, c_size_t, _P]),
    'caspr_linear_wgrad_workspace_bytes': (c_size_t, [c_longlong, c_int, c_int]),
    'caspr_linear_wgrad': (c_int, [_P, c_int, _P, c_int, c_longlong, c_int, c_int, c_int, _P, _P, _P, c_size_t, _P]),
    'caspr_linear_wgrad_tc_workspace_bytes': (c_size_t, [c_longlong, c_int, c_int]),
    'caspr_linear_wgrad_tc': (c_int, [_P, c_int, _P, c_int, c_longlong, c_int, c_int, c_int, _P, _P, _P, _P, c_size_t, _P]),
    'caspr_colsum_workspace_bytes': (c_size_t, [c_longlong, c_int]),
    'caspr_colsum': (c_int, [_P, c_int, c_longlong, c_int, _P, c_int, _P, c_size_t, _P]),
    'caspr_group_points_bwd': (c_int, [_P, c_int, _P, c_int, c_int, c_int, c_int, c_int, _P, c_int, _P]),
    'caspr_three_interp_bwd': (c_int, [_P, c_int, _P, _P, c_int, c_int, c_int, c_int, _P, c_int, _P]),
    'caspr_rows_update': (c_int, [_P, c_int, c_longlong, c_int, c_int, _P, c_int, _P, c_int, _P]),
    'caspr_transpose': (c_int, [_P, c_int, c_int, _P, _P]),
    'caspr_assemble_batch': (c_int, [_P, _P, _P, _P, c_int, c_int, c_int, _P, c_int, _P, c_int, c_int, c_double, c_int,
                                     _P, _P, _P]),
    'caspr_chamfer': (c_int, [_P, _P, c_int, c_int, c_int, _P, _P, _P]),
    'caspr_tnocs_error': (c_int, [_P, _P, c_int, c_int, _P, _P, _P]),
    'caspr_emd_workspace_bytes': (c_size_t, [c_int, c_int, c_int]),
    'caspr_emd': (c_int, [_P, _P, c_int, c_int, c_int, _P, _P, c_size_t, _P]),
    'caspr_sa_mlp_tc_supported': (c_int, [c_int, c_int, c_int, c_int, c_int]),
    'caspr_sa_mlp_tc_workspace_bytes': (c_size_t, [c_longlong, c_int, c_int, c_int]),
    'caspr_sa_mlp_tc': (c_int, [_P, c_int, c_longlong, c_int, c_int] + [_P, _P, _P, _P, c_int] * 3 +
                        [c_float, _P, c_int, _P, c_size_t, _P]),
    'caspr_sa_mlp_tc_grouped': (c_int, [_P, _P, _P, c_int, c_int, _P, c_int, c_int, c_int, c_int] + [_P, _P, _P, _P, c_int] * 3 +
                                [c_float, _P, c_int, _P, c_size_t, _P]),
    'caspr_sa_mlp_tc_delayed_workspace_bytes': (c_size_t, [c_longlong, c_int, c_int]),
    'caspr_sa_mlp_tc_delayed': (c_int, [_P, _P, _P, c_int, _P, c_int, c_int, c_int, c_int, _P, c_int, _P, _P, _P, c_int] +
                                [_P, _P, _P, _P, c_int] * 2 + [c_float, _P, c_int, _P, c_size_t, _P]),
    'caspr_cnf_fused_debug_read': (c_int, [_P, c_int]),
    'caspr_ransac_pose_workspace_bytes': (c_size_t, [c_int, c_int]),
    'caspr_ransac_pose': (c_int, [_P, _P, _P, c_int, c_int, c_int, c_float, c_int, _P, _P, _P, _P, _P, _P, _P, c_size_t, _P]),
}

if not os.path.exists(LIB_PATH):
    raise ImportError(
        'libcaspr_b200.so not found at %s.  caspr_b200 has no CPU or PyTorch fallback: build the '
        'CUDA library first with `python -m caspr_b200.build`.' % LIB_PATH)

lib = ctypes.CDLL(LIB_PATH)
for _name, (_res, _args) in SIGNATURES.items():
    _fn = getattr(lib, _name)          # AttributeError here = header / library out of sync
    _fn.restype = _res
    _fn.argtypes = _args


def status_string(status):
    return lib.caspr_status_string(int(status)).decode()


def check(status, where):
    if status != 0:
        raise CasprError(status, where)
