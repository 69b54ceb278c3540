"""AZ.Save / AZ.Load container (agogo.go:175-209): the gob stream writer/reader of agogo_b200/gobfmt.py.  The wire
primitives are checked against the byte sequences the encoding/gob documentation spells out; the container round-trips
bit-exactly; host.AZ.Save/Load through it restore both agents' nets (Load gives A and B the stored net and clears
useDummy).  The tensor.Dense field layout inside is a restatement from memory — unverified against a Go build."""
import numpy as np
import pytest

from agogo_b200 import _capi as K
from agogo_b200 import gobfmt as G
from agogo_b200 import host


def test_gob_primitives_match_the_documented_examples():
    # encoding/gob doc: "7 is transmitted as 07", "256 is transmitted as (FE 01 00)"
    assert G.enc_uint(7) == bytes([7]) and G.enc_uint(256) == bytes([0xFE, 0x01, 0x00])
    # doc example: struct { A, B int } {7, 8} ... the int 7 is sent as 0e (7 << 1), -129 as (FE 01 01)
    assert G.enc_int(7) == bytes([0x0E]) and G.enc_int(-129) == bytes([0xFE, 0x01, 0x01])
    # doc: float 17.0 = 0x4031000000000000, byte-reversed 0x3140 -> FE 31 40
    assert G.enc_float(17.0) == bytes([0xFE, 0x31, 0x40])
    assert G.enc_string("hello") == bytes([5]) + b"hello"
    r = G.Reader(G.enc_uint(300) + G.enc_int(-5) + G.enc_float(-0.375) + G.enc_int(1 << 40))
    assert r.uint() == 300 and r.int() == -5 and r.float() == -0.375 and r.int() == 1 << 40 and r.eof()
    # first user type id is 65: a type definition message starts with its negated id, FF 81
    e = G.Encoder()
    e.value_slice("[]int", G.T_INT, [1, 2, 3], G.enc_int)
    b = e.bytes()
    assert b[1:3] == bytes([0xFF, 0x81])


def test_type_descriptor_matches_the_documented_point_example():
    """The one complete byte dump in the encoding/gob documentation: `type Point struct { X, Y int }`, value {22, 33}.  Our
    stream writer's type-definition layer (message framing, negated id, wireType / CommonType nesting, field deltas) must
    reproduce the 32 bytes of the descriptor and the 8 bytes of the value; the structType field list is spelled here, the
    rest comes from the same `_define` every checkpoint type goes through."""
    doc_type = bytes.fromhex("1f ff 81 03 01 01 05 50 6f 69 6e 74 01 ff 82 00 01 02 01 01 58 01 04 00 01 01 59 01 04 00 00 00".replace(" ", ""))
    doc_value = bytes.fromhex("07 ff 82 01 2c 01 42 00".replace(" ", ""))
    e = G.Encoder()
    fields = G.enc_uint(1) + G.enc_uint(2)                                      # structType.Field (field 1), two entries
    for nm in ("X", "Y"):
        fields += G.enc_uint(1) + G.enc_string(nm) + G.enc_uint(1) + G.enc_int(G.T_INT) + G.enc_uint(0)   # fieldType{Name, Id}
    tid = e._define(("struct", "Point"), "Point", G.WT_STRUCT, fields)
    assert tid == 65 and e.bytes() == doc_type
    e._message(G.enc_int(tid) + G.enc_uint(1) + G.enc_int(22) + G.enc_uint(1) + G.enc_int(33) + G.enc_uint(0))
    assert e.bytes() == doc_type + doc_value


def test_vectorised_float_encoding_equals_scalar():
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.normal(size=500).astype(np.float32), np.array([0.0, -0.0, 1.0, 2.0, -2.0, 0.5, 1e-30, 3e38, np.float32(1) / 3], np.float32)])
    assert G.enc_floats(x) == b"".join(G.enc_float(float(v)) for v in x)


def test_container_round_trip_bit_exact():
    rng = np.random.default_rng(1)
    tensors = [rng.normal(size=s).astype(np.float32) for s in [(3, 2, 3, 3), (3, 3, 3, 3), (18, 10), (4, 10), (4, 1)]]
    tensors[0][0, 0, 0, 0] = 0.0
    blob = G.save_stream(tensors)
    back = G.load_stream(blob)
    assert len(back) == len(tensors)
    for a, b in zip(tensors, back):
        assert a.shape == b.shape and (a.view(np.uint32) == b.view(np.uint32)).all()
    with pytest.raises(ValueError):
        G.load_stream(blob[:len(blob) // 2])


def test_az_save_load_gob(oracle, tmp_path):
    from tests.test_host_learn import _c1_conf
    conf = _c1_conf(batch=20, sims=10)
    az = host.AZ(host.Game(K.GAME_MNK, 3, 3, 3), conf, lib=oracle, n_games=4, seed=3)
    f = str(tmp_path / "ttt.model")
    az.Save(f)
    want = az.engine.net_get(0)
    az2 = host.AZ(host.Game(K.GAME_MNK, 3, 3, 3), conf, lib=oracle, n_games=4, seed=99)
    assert not (az2.engine.net_get(0) == want).all()
    az2.Load(f)
    assert (az2.engine.net_get(0).view(np.uint32) == want.view(np.uint32)).all()
    assert (az2.engine.net_get(1).view(np.uint32) == want.view(np.uint32)).all()   # agogo.go:196-206: B gets the same net
    assert az2.useDummy is False
    # the .npz side format still works
    az.Save(str(tmp_path / "ttt.npz")); az2.Load(str(tmp_path / "ttt.npz"))
