"""CPU: the C-ABI shared library loads and exports every symbol include/b200rl.h declares (no compute calls)."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    src = open(os.path.join(ROOT, "include", "b200rl.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(b200rl_[a-z0-9_]+)\s*\(", src)))


def test_header_and_binding_agree():
    from rllab_b200 import _lib
    assert _header_functions() == sorted(_lib.SIGNATURES.keys())


def test_library_loads_and_exports_every_symbol():
    from rllab_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    lib = _lib.load()
    for name in _header_functions():
        assert hasattr(lib, name), name
    assert lib.b200rl_version() == 100
    assert lib.b200rl_ws_doubles() > 0


def test_unsupported_shapes_fail_loudly():
    from rllab_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        pytest.skip("library not built")
    with pytest.raises(_lib.B200RLError):
        _lib.policy_num_params(7, 32, 32, 2)            # (O, A) pair not compiled in
    with pytest.raises(_lib.B200RLError):
        _lib.policy_num_params(4, 16, 16, 1)            # hidden size not compiled in
    assert _lib.policy_num_params(4, 32, 32, 1) == 1250
    assert _lib.policy_num_params(3, 32, 32, 1) == 1218
    info = _lib.env_info(_lib.ENV_CARTPOLE)
    assert (info["obs_dim"], info["act_dim"], info["lb"], info["ub"]) == (4, 1, [-10.0], [10.0])
    with pytest.raises(_lib.B200RLError):
        _lib.env_info(99)
