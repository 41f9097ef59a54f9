"""CPU: libpaacb.so loads, exports exactly what include/paacb.h declares, and refuses to run without a GPU."""
import ctypes as C
import os
import re

import pytest
import torch

from paac_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def lib():
    from paac_b200 import build
    build.build()            # incremental: a no-op when libpaacb.so is newer than its sources
    return _lib.load()


def declared_symbols():
    hdr = open(os.path.join(ROOT, 'include', 'paacb.h')).read()
    hdr = re.sub(r'/\*.*?\*/', '', hdr, flags=re.S)
    return sorted(set(re.findall(r'\b(paacb_[a-z0-9_]+)\s*\(', hdr)))


def test_every_declared_symbol_is_exported_and_bound(lib):
    syms = declared_symbols()
    assert len(syms) >= 23
    for s in syms:
        assert hasattr(lib, s), s
    assert sorted(_lib.PROTOTYPES) == syms        # the ctypes table covers the header exactly


def test_version_and_error_channel(lib):
    assert lib.paacb_version() == 102
    h = C.c_void_p()
    rc = lib.paacb_create(C.byref(h), 7, 6, 0)    # bad arch -> EINVAL, message set, no exception
    assert rc == -1 and b'arch' in lib.paacb_last_error()
    rc = lib.paacb_create(C.byref(h), _lib.ARCH_NATURE, 99, 0)
    assert rc == -1 and b'num_actions' in lib.paacb_last_error()


@pytest.mark.skipif(torch.cuda.is_available(), reason='CPU-only behaviour')
def test_fails_loudly_without_gpu(lib):
    h = C.c_void_p()
    rc = lib.paacb_create(C.byref(h), _lib.ARCH_NATURE, 6, 0)
    assert rc == -2 and b'no CPU path' in lib.paacb_last_error()
    from paac_b200.policy_v_network import NaturePolicyVNetwork
    conf = dict(name='x', num_actions=6, clip_norm=3.0, clip_norm_type='global', device='/gpu:0',
                entropy_regularisation_strength=0.02)
    with pytest.raises(_lib.PaacbError):
        NaturePolicyVNetwork(conf)


def test_sass_is_sm100(lib):
    import subprocess
    out = subprocess.run(['cuobjdump', '-lelf', _lib.LIB_PATH], stdout=subprocess.PIPE, text=True).stdout
    assert 'sm_100a' in out and 'sm_90' not in out and 'sm_80' not in out
