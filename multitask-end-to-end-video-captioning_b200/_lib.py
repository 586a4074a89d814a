"""ctypes binding of libs2vt_b200.so (the C ABI of include/s2vt.h).

There is no CPU fallback: if the shared library is missing or a symbol cannot be resolved, importing this
module raises.  Build it with `python -m __graft_entry__` / `__graft_entry__.build()`.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('S2VT_LIB') or os.path.join(HERE, 'libs2vt_b200.so')   # S2VT_LIB: experimental builds (debug only)


class S2vtConfig(C.Structure):
    _fields_ = [('dim_image', C.c_int32), ('word_dim', C.c_int32), ('lstm_dim', C.c_int32), ('n_words', C.c_int32),
                ('n_video_steps', C.c_int32), ('n_caption_steps', C.c_int32), ('n_attributes', C.c_int32),
                ('precision', C.c_int32), ('gemm_backend', C.c_int32), ('dropout_keep', C.c_float)]


class S2vtAttConfig(C.Structure):
    _fields_ = [('dim_image', C.c_int32), ('dim_hidden', C.c_int32), ('n_words', C.c_int32), ('n_video_steps', C.c_int32),
                ('n_caption_steps', C.c_int32), ('precision', C.c_int32), ('dropout_keep', C.c_float), ('hinge_beta', C.c_float),
                ('hinge_m', C.c_float), ('reg_frames', C.c_int32)]


PREC_BF16, PREC_FP32 = 0, 1
GEMM_AUTO, GEMM_MMA_SYNC, GEMM_TCGEN05 = 0, 1, 2
S2VT_OK, S2VT_EINVAL, S2VT_ENOTFOUND, S2VT_ESHAPE, S2VT_ECUDA, S2VT_ENOSPACE, S2VT_ESTATE = 0, -1, -2, -3, -4, -5, -6

_vp, _i32, _i64, _u32, _u64, _f32, _sz = C.c_void_p, C.c_int32, C.c_int64, C.c_uint32, C.c_uint64, C.c_float, C.c_size_t

# name -> (restype, argtypes); every symbol include/s2vt.h declares
SIGNATURES = {
    's2vt_create': (_i32, [C.POINTER(S2vtConfig), C.POINTER(_vp)]),
    's2vt_destroy': (None, [_vp]),
    's2vt_last_error': (C.c_char_p, [_vp]),
    's2vt_num_params': (_sz, [_vp]),
    's2vt_state_bytes': (_sz, [_vp]),
    's2vt_workspace_bytes': (_sz, [_vp, _i32, _i32, _i32]),
    's2vt_bind': (_i32, [_vp, _vp, _sz, _vp, _sz]),
    's2vt_params': (_vp, [_vp]),
    's2vt_grads': (_vp, [_vp]),
    's2vt_adam_m': (_vp, [_vp]),
    's2vt_adam_v': (_vp, [_vp]),
    's2vt_num_variables': (_i32, [_vp]),
    's2vt_variable_info': (_i32, [_vp, _i32, C.POINTER(C.c_char_p), C.POINTER(_i64), C.POINTER(_i64 * 2), C.POINTER(_i32)]),
    's2vt_load_param': (_i32, [_vp, C.c_char_p, _vp, C.POINTER(_i64), _i32, _vp]),
    's2vt_refresh': (_i32, [_vp, _vp]),
    's2vt_set_overlap': (_i32, [_vp, _i32]),
    's2vt_set_reuse_frontend': (_i32, [_vp, _i32]),
    's2vt_greedy': (_i32, [_vp, _vp, _i32, _vp, _vp]),
    's2vt_rollout': (_i32, [_vp, _vp, _i32, _i32, _u64, _u32, _vp, _vp, _vp]),
    's2vt_caption_masks': (_i32, [_vp, _vp, _i32, _vp, _vp, _vp]),
    's2vt_teacher_forward': (_i32, [_vp, _vp, _i32, _vp, _i32, _u64, _u32, _vp, _vp, _vp]),
    's2vt_rl_backward': (_i32, [_vp, _vp, _i32, _vp, _vp, _vp, _vp, _i32, _f32, _f32, _i32, _u64, _u32, _vp, _vp]),
    's2vt_xe_backward': (_i32, [_vp, _vp, _i32, _vp, _vp, _i32, _f32, _f32, _f32, _f32, _i32, _u64, _u32, _vp, _vp]),
    's2vt_xe_backward_sharded': (_i32, [_vp, _vp, _i32, _vp, _vp, _i32, _f32, _f32, _f32, _vp, _i32, _f32, _i32, _u64, _u32, _vp, _vp]),
    's2vt_debug_gemm': (_i32, [_vp, _i32, _i32, _i32, _i32, _i32, _vp]),
    's2vt_attribute_backward': (_i32, [_vp, _vp, _i32, _vp, _f32, _vp, _vp]),
    's2vt_optimizer_step': (_i32, [_vp, _f32, _f32, _i64, _i32, _vp, _vp]),
    's2vt_grad_segment_ready': (_i32, [_vp, _i32, _vp, C.POINTER(_i64), C.POINTER(_i64)]),
    's2vt_peer_export': (_i32, [_vp, _vp, C.POINTER(_i64), _vp]),
    's2vt_peer_connect': (_i32, [_vp, _i32, _i32, _vp, _vp, _vp]),
    's2vt_peer_allreduce': (_i32, [_vp, _vp]),
    's2vt_peer_optimizer_step': (_i32, [_vp, _f32, _f32, _i64, _i32, _vp, _vp]),
    's2vt_peer_gather_state': (_i32, [_vp, _vp]),
    's2vt_peer_disconnect': (_i32, [_vp]),
    's2vt_launch_count': (C.c_longlong, [_vp]),
    's2vt_profile': (_i32, [_vp, _i32]),
    's2vt_profile_read': (_i32, [_vp, _vp, _vp, _vp]),
    's2vt_debug_probe': (_i32, [_vp]),
    's2vt_profile_shapes': (_i32, [_vp, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    's2vt_beam_search': (_i32, [_vp, _vp, _i32, _i32, _f32, _vp, _vp, _vp, _vp, _vp]),
    's2vt_beam_init': (_i32, [_vp, _vp, _vp, _vp, _vp]),
    's2vt_beam_step': (_i32, [_vp, _vp, _vp, _vp, _i32, _vp, _vp, _vp, _vp, _vp]),
    'ciderd_corpus_create': (_i32, [_vp, _vp, _i64, _vp, _i64, _vp, C.POINTER(_vp)]),
    'ciderd_corpus_destroy': (None, [_vp]),
    'ciderd_corpus_device_bytes': (_sz, [_vp]),
    'ciderd_corpus_serialize': (_i32, [_vp, _vp]),
    'ciderd_score': (_i32, [_vp, _vp, _vp, _i32, _i32, _vp, _vp, _vp]),
    's2vt_reward_corpus_create': (_i32, [_vp, _vp, _i64, _vp, _i64, C.POINTER(_vp)]),
    's2vt_reward_corpus_destroy': (None, [_vp]),
    's2vt_reward_corpus_device_bytes': (_sz, [_vp]),
    's2vt_reward_corpus_serialize': (_i32, [_vp, _vp]),
    's2vt_bleu_score': (_i32, [_vp, _vp, _vp, _i32, _i32, _vp, _vp, _vp]),
    's2vt_rouge_score': (_i32, [_vp, _vp, _vp, _i32, _i32, _i32, _vp, _vp]),
    's2vt_att_create': (_i32, [C.POINTER(S2vtAttConfig), C.POINTER(_vp)]),
    's2vt_att_destroy': (None, [_vp]),
    's2vt_att_last_error': (C.c_char_p, [_vp]),
    's2vt_att_num_params': (_sz, [_vp]),
    's2vt_att_state_bytes': (_sz, [_vp]),
    's2vt_att_workspace_bytes': (_sz, [_vp, _i32, _i32]),
    's2vt_att_bind': (_i32, [_vp, _vp, _sz, _vp, _sz]),
    's2vt_att_params': (_vp, [_vp]),
    's2vt_att_num_variables': (_i32, [_vp]),
    's2vt_att_variable_info': (_i32, [_vp, _i32, C.POINTER(C.c_char_p), C.POINTER(_i64), C.POINTER(_i64 * 2), C.POINTER(_i32)]),
    's2vt_att_load_param': (_i32, [_vp, C.c_char_p, _vp, C.POINTER(_i64), _i32, _vp]),
    's2vt_att_refresh': (_i32, [_vp, _vp]),
    's2vt_att_greedy': (_i32, [_vp, _vp, _i32, _vp, _vp, _vp]),
    's2vt_att_xe_loss': (_i32, [_vp, _vp, _i32, _vp, _vp, _u64, _u32, _vp, _vp, _vp]),
    's2vt_att_grads': (_vp, [_vp]),
    's2vt_att_adam_m': (_vp, [_vp]),
    's2vt_att_adam_v': (_vp, [_vp]),
    's2vt_att_xe_backward': (_i32, [_vp, _vp, _i32, _vp, _vp, _u64, _u32, _vp, _vp]),
    's2vt_att_optimizer_step': (_i32, [_vp, _f32, _f32, _i64, _vp, _vp]),
    's2vt_att_launch_count': (C.c_longlong, [_vp]),
    # include/s2vt_io.h (host only)
    's2vt_io_last_error': (C.c_char_p, []),
    's2vt_features_open': (_i32, [C.c_char_p, _i32, C.POINTER(_vp)]),
    's2vt_features_open_memory': (_i32, [_vp, _sz, _i32, C.POINTER(_vp)]),
    's2vt_features_close': (None, [_vp]),
    's2vt_features_num_videos': (_i64, [_vp]),
    's2vt_features_num_frames': (_i32, [_vp]),
    's2vt_features_dim': (_i32, [_vp]),
    's2vt_features_video_id': (C.c_char_p, [_vp, _i64]),
    's2vt_features_find': (_i64, [_vp, C.c_char_p]),
    's2vt_features_read': (_i32, [_vp, _vp, _i64, _vp, _i32]),
    's2vt_ckpt_open': (_i32, [C.c_char_p, C.POINTER(_vp)]),
    's2vt_ckpt_close': (None, [_vp]),
    's2vt_ckpt_format': (_i32, [_vp]),
    's2vt_ckpt_num_tensors': (_i32, [_vp]),
    's2vt_ckpt_tensor_info': (_i32, [_vp, _i32, C.POINTER(C.c_char_p), C.POINTER(_i32), C.POINTER(_i32), C.POINTER(_i64 * 8)]),
    's2vt_ckpt_find': (_i32, [_vp, C.c_char_p]),
    's2vt_ckpt_read_f32': (_i32, [_vp, _i32, _vp, _i64]),
}

_lib = None


def load():
    """Load the shared library once and attach the prototypes.  Raises if it is absent -- never falls back."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError('%s is missing: run `python -c "import __graft_entry__ as g; g.build()"` first '
                          '(there is no CPU fallback)' % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)        # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class S2vtError(RuntimeError):
    def __init__(self, code, msg):
        RuntimeError.__init__(self, 's2vt error %d: %s' % (code, msg))
        self.code = code


def check_io(code):
    """Return-code check for the host-only entry points of include/s2vt_io.h."""
    if code != 0:
        msg = load().s2vt_io_last_error()
        raise S2vtError(code, msg.decode() if msg else '')


def check(handle, code):
    if code != 0:
        msg = load().s2vt_last_error(handle)
        raise S2vtError(code, msg.decode() if msg else '')
