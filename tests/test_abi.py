"""CPU checks of the C-ABI boundary: the library builds/loads, exports every symbol include/*.h declares, and the
host-only entry points (create / sizing / variable table) behave.  No compute is launched."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def lib():
    import __graft_entry__ as ge
    ge.build()
    import s2vt_b200
    return s2vt_b200._lib.load()


def declared_functions():
    src = ''
    for h in sorted(os.listdir(os.path.join(ROOT, 'include'))):
        if h.endswith('.h'):
            src += open(os.path.join(ROOT, 'include', h)).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    names = re.findall(r'\b((?:s2vt|ciderd)_[a-z0-9_]+)\s*\(', src)
    return sorted(set(names))


def test_header_symbols_are_exported_and_bound(lib):
    import s2vt_b200
    names = declared_functions()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), 'library does not export %s' % n
        assert n in s2vt_b200._lib.SIGNATURES, 'ctypes binding misses %s' % n
    assert sorted(s2vt_b200._lib.SIGNATURES) == names


def test_create_sizes_and_variable_table(lib):
    import s2vt_b200
    cfg = s2vt_b200._lib.S2vtConfig(1536, 500, 1000, 9972, 5, 35, 0, 0, 0, 0.9)
    h = C.c_void_p()
    assert lib.s2vt_create(C.byref(cfg), C.byref(h)) == 0
    assert lib.s2vt_num_params(h) == 31744472        # 31 744 472 parameters
    assert lib.s2vt_num_variables(h) == 9
    seen = {}
    for i in range(9):
        name, off, shape, nd = C.c_char_p(), C.c_int64(), (C.c_int64 * 2)(), C.c_int32()
        assert lib.s2vt_variable_info(h, i, C.byref(name), C.byref(off), C.byref(shape), C.byref(nd)) == 0
        seen[name.value.decode()] = (shape[0], shape[1], nd.value)
    assert seen['Wemb'] == (9972, 500, 2) and seen['embed_word_W'] == (1000, 9972, 2) and seen['encode_image_b'] == (500, 0, 1)
    assert seen['s2vt/LSTM1/basic_lstm_cell/weights'] == (1500, 4000, 2)
    assert seen['s2vt/LSTM2/basic_lstm_cell/weights'] == (2500, 4000, 2)
    assert lib.s2vt_state_bytes(h) > 4 * 4 * 31744472
    small, big = lib.s2vt_workspace_bytes(h, 8, 40, 3), lib.s2vt_workspace_bytes(h, 64, 320, 5)
    assert 0 < small < big < 40 * 2 ** 30
    lib.s2vt_destroy(h)


def test_create_rejects_bad_config(lib):
    import s2vt_b200
    h = C.c_void_p()
    for bad in (s2vt_b200._lib.S2vtConfig(0, 500, 1000, 9972, 5, 35, 0, 0, 0, 0.9),
                s2vt_b200._lib.S2vtConfig(1536, 500, 1000, 70000, 5, 35, 0, 0, 0, 0.9),
                s2vt_b200._lib.S2vtConfig(1536, 500, 1000, 9972, 5, 35, 0, 7, 0, 0.9)):
        assert lib.s2vt_create(C.byref(bad), C.byref(h)) == s2vt_b200._lib.S2VT_EINVAL


def test_unbound_handle_fails_loudly(lib):
    import s2vt_b200
    cfg = s2vt_b200._lib.S2vtConfig(64, 16, 16, 50, 2, 4, 0, 1, 0, 1.0)
    h = C.c_void_p()
    assert lib.s2vt_create(C.byref(cfg), C.byref(h)) == 0
    assert lib.s2vt_refresh(h, None) == s2vt_b200._lib.S2VT_ESTATE
    assert lib.s2vt_greedy(h, None, 1, C.c_void_p(8), None) != 0
    assert b'bind' in lib.s2vt_last_error(h)
    lib.s2vt_destroy(h)


def test_model_refuses_to_run_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip('CUDA present')
    import s2vt_b200
    with pytest.raises(RuntimeError):
        s2vt_b200.Video_Caption_Generator()
