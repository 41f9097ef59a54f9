"""CPU: the TensorFlow-1 checkpoint-bundle reader/writer (paac_b200/tf_bundle.py) pinned against what the reference ships.

tests/golden/tf_index_*.index are byte copies of the reference's own ``pretrained/<game>/checkpoints/-80000000.index``
files (data artefacts, produced by TF 1.0.1's Saver; copied by oracle/make_golden.py).  The reader must recover the
variable names / shapes the shipped training graph declares, and the writer must re-serialise the table BYTE FOR BYTE."""
import glob
import json
import os

import numpy as np
import pytest

from paac_b200 import tf_bundle as tb

GOLD = os.path.join(os.path.dirname(__file__), 'golden')
INDEXES = sorted(glob.glob(os.path.join(GOLD, 'tf_index_*.index')))


def test_crc32c_known_answers():
    assert tb.crc32c(b'123456789') == 0xE3069283            # the standard CRC-32C check value
    assert tb.crc32c(b'') == 0
    assert tb.crc32c(bytes(32)) == 0x8A9136AA               # RFC 3720 B.4: 32 bytes of zeros
    assert tb.crc32c(bytes([0xff] * 32)) == 0x62A8AB43      # RFC 3720 B.4: 32 bytes of ones
    rng = np.random.RandomState(0)
    big = rng.bytes(200000 + 13)                            # the chunked NumPy path == the byte loop, also when chained
    assert tb.crc32c(big) == tb._crc32c_bytes(big)
    assert tb.crc32c(big[70001:], tb.crc32c(big[:70001])) == tb._crc32c_bytes(big)
    assert tb.unmask_crc(tb.mask_crc(0x12345678)) == 0x12345678


@pytest.mark.parametrize('path', INDEXES, ids=[os.path.basename(p) for p in INDEXES])
def test_shipped_index_round_trips_byte_for_byte(path):
    header, entries = tb.read_index(path)
    assert header == tb.HEADER_VALUE
    assert tb.serialize_index(entries) == open(path, 'rb').read()
    # 10 variables x (value, rms slot, momentum slot), float32, packed back to back in key order
    assert len(entries) == 30
    offset = 0
    for name in sorted(entries):
        e = entries[name]
        assert e.dtype == tb.DT_FLOAT and e.shard_id == 0 and e.offset == offset
        assert e.size == 4 * int(np.prod(e.shape))
        offset += e.size


def test_shipped_index_matches_the_training_graph_pins():
    assert len(INDEXES) >= 7
    pins = json.load(open(os.path.join(GOLD, 'tf_graph_pins.json')))
    for path in INDEXES:
        game = os.path.basename(path)[len('tf_index_'):-len('.index')]
        _, entries = tb.read_index(path)
        want = pins[game]['shapes'] if game in pins else None
        names = [n for n in entries if 'OptimizerVariables' not in n]
        assert len(names) == 10
        assert entries['local_learning_1/conv1_weights'].shape == (8, 8, 4, 16)
        assert entries['local_learning_1/fc3_weights'].shape == (2592, 256)
        a = entries['local_learning_2/actor_output_weights'].shape[1]
        assert entries['local_learning_2/critic_output_weights'].shape == (256, 1)
        if want is not None:
            for n in names:
                assert tuple(want[n]) == entries[n].shape, (game, n)
        assert 4 <= a <= 18


def test_write_then_read(tmp_path):
    rng = np.random.RandomState(1)
    tensors = {'local_learning_1/conv1_weights': rng.randn(8, 8, 4, 32).astype(np.float32),
               'local_learning_1/conv1_biases': rng.randn(32).astype(np.float32),
               'local_learning_1/fc4_weights': rng.randn(3136, 512).astype(np.float32),
               'local_learning_2/critic_output_biases': rng.randn(1).astype(np.float32)}
    prefix = str(tmp_path / 'checkpoints' / '-1234')
    tb.write_bundle(prefix, tensors)
    back = tb.read_bundle(prefix)
    assert list(back) == sorted(tensors)
    for k in tensors:
        assert np.array_equal(back[k], tensors[k])
    # corrupt one byte of the data file: the per-tensor checksum catches it
    data = bytearray(open(prefix + '.data-00000-of-00001', 'rb').read())
    data[100] ^= 1
    open(prefix + '.data-00000-of-00001', 'wb').write(bytes(data))
    with pytest.raises(ValueError):
        tb.read_bundle(prefix)


def test_saver_layout_and_rotation(tmp_path):
    import torch
    from paac_b200.session import Saver
    state = {'conv1_weights': torch.randn(8, 8, 4, 16), 'actor_output_biases': torch.randn(6),
             'conv1_weights/OptimizerVariables': torch.ones(8, 8, 4, 16)}
    loaded = {}
    saver = Saver(lambda: state, loaded.update, max_to_keep=2)
    folder = str(tmp_path / 'checkpoints')
    for step in (100, 200, 300):
        saver.save(None, folder, step)
    assert sorted(os.listdir(folder)) == ['-200.data-00000-of-00001', '-200.index', '-300.data-00000-of-00001', '-300.index',
                                          'checkpoint']
    latest = Saver.latest_checkpoint(folder)
    assert latest.endswith('-300') and int(latest[latest.rindex('-') + 1:].split('.')[0]) == 300     # networks.py:134
    _, entries = tb.read_index(latest + '.index')
    assert set(entries) == {'local_learning_1/conv1_weights', 'local_learning_2/actor_output_biases',
                            'local_learning_1/conv1_weights/OptimizerVariables'}
    saver.restore(None, latest)
    for k in state:
        assert torch.equal(loaded[k], state[k])
