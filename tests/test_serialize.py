"""Actor / module wire format (relearn_b200/serialize.py) against the reference's own known-answer tests:
the serde_test token streams of src/torch/serialize.rs:176-340 are transcribed below; CBOR / JSON encodings are
checked against bytes assembled by hand from RFC 8949 / what serde_json prints."""
import numpy as np
import pytest

from relearn_b200 import serialize as S
from relearn_b200.modules import init_params


def _tokens(kind, shape, requires_grad, data):
    return ([("Struct", "TensorDef", 5), ("Str", "kind"), ("UnitVariant", "KindDef", kind), ("Str", "shape"), ("Seq", len(shape))]
            + [("I64", n) for n in shape]
            + [("SeqEnd",), ("Str", "requires_grad"), ("Bool", requires_grad), ("Str", "byte_order"),
               ("UnitVariant", "ByteOrder", "LittleEndian"), ("Str", "data"), ("BorrowedBytes", data), ("StructEnd",)])


def test_ser_de_tokens_0d_i32_tensor():
    # serialize.rs:176-221: Tensor::of_slice(&[0x1234_5678_i32]).reshape(&[])
    d = S.tensor_def(np.array(0x12345678, np.int32))
    assert S.tensor_def_tokens(d) == _tokens("Int", [], False, bytes([0x78, 0x56, 0x34, 0x12]))
    t, rg = S.tensor_from_def(d)
    assert t.shape == () and t.dtype == np.int32 and int(t) == 0x12345678 and rg is False


def test_ser_de_tokens_empty_f32_tensor():
    # serialize.rs:223-257
    d = S.tensor_def(np.zeros((0,), np.float32))
    assert S.tensor_def_tokens(d) == _tokens("Float", [0], False, b"")
    assert S.tensor_from_def(d)[0].shape == (0,)


def test_ser_de_tokens_1d_f32_tensor_requires_grad():
    # serialize.rs:259-293: bytes [0, 0, 128, 63]
    d = S.tensor_def(np.array([1.0], np.float32), requires_grad=True)
    assert S.tensor_def_tokens(d) == _tokens("Float", [1], True, bytes([0, 0, 128, 63]))
    t, rg = S.tensor_from_def(d)
    assert rg is True and t.tolist() == [1.0]


def test_ser_de_tokens_2d_u8_tensor_and_roundtrip():
    # serialize.rs:295-337 and to_from_tensordef :339-345
    t0 = np.array([1, 2, 3, 4, 5, 6], np.uint8).reshape(2, 3)
    d = S.tensor_def(t0)
    assert S.tensor_def_tokens(d) == _tokens("Uint8", [2, 3], False, bytes([1, 2, 3, 4, 5, 6]))
    assert np.array_equal(S.tensor_from_def(d)[0], t0)


def test_tensor_def_cbor_bytes_by_hand():
    """map(5){ "kind":"Float", "shape":[1], "requires_grad":true, "byte_order":"LittleEndian", "data":h'0000803f' }"""
    d = S.tensor_def(np.array([1.0], np.float32), requires_grad=True)
    want = (b"\xa5" + b"\x64kind" + b"\x65Float" + b"\x65shape" + b"\x81\x01" + b"\x6drequires_grad" + b"\xf5"
            + b"\x6abyte_order" + b"\x6cLittleEndian" + b"\x64data" + b"\x44\x00\x00\x80\x3f")
    assert S.to_cbor(d) == want
    assert S.from_cbor(want) == d
    assert S.to_json(d) == '{"kind":"Float","shape":[1],"requires_grad":true,"byte_order":"LittleEndian","data":[0,0,128,63]}'
    back = S.from_json(S.to_json(d))
    assert np.array_equal(S.tensor_from_def(back)[0], np.array([1.0], np.float32))


def test_cbor_scalars_follow_rfc8949_examples():
    # RFC 8949 appendix A
    for v, hexs in [(0, "00"), (23, "17"), (24, "1818"), (100, "1864"), (1000, "1903e8"), (1000000, "1a000f4240"),
                    (1000000000000, "1b000000e8d4a51000"), (-1, "20"), (-100, "3863"), (-1000, "3903e7"),
                    (1.0, "f93c00"), (1.5, "f93e00"), (100000.0, "fa47c35000"), (1.1, "fb3ff199999999999a"),
                    (float("inf"), "f97c00"), (float("-inf"), "f9fc00"), (False, "f4"), (True, "f5"), (None, "f6"),
                    (b"\x01\x02\x03\x04", "4401020304"), ("IETF", "6449455446"), ([1, [2, 3], [4, 5]], "8301820203820405")]:
        assert S.to_cbor(v).hex() == hexs, v
        got = S.from_cbor(bytes.fromhex(hexs))
        assert got == v and type(got) is type(v)
    # the decoder takes every width and indefinite lengths (a Rust writer may choose differently)
    assert S.from_cbor(bytes.fromhex("fb3ff0000000000000")) == 1.0
    assert S.from_cbor(bytes.fromhex("9f018202039f0405ffff")) == [1, [2, 3], [4, 5]]
    assert S.from_cbor(bytes.fromhex("bf61610161629f0203ffff")) == {"a": 1, "b": [2, 3]}
    assert S.from_cbor(bytes.fromhex("5f42010243030405ff")) == b"\x01\x02\x03\x04\x05"
    with pytest.raises(ValueError):
        S.from_cbor(bytes.fromhex("8301"))
    with pytest.raises(ValueError):
        S.from_cbor(bytes.fromhex("0000"))


def test_mlp_and_policy_actor_roundtrip(tmp_path):
    params = init_params(np.random.default_rng(3), 5, 128, 2)
    mlp = S.mlp_to_serde(params, 5, [128], 2)
    assert list(mlp) == ["layers", "activation", "output_activation"] and mlp["activation"] == "Relu"
    assert [list(layer) for layer in mlp["layers"]] == [["kernel", "bias"]] * 2
    assert mlp["layers"][0]["kernel"]["shape"] == [128, 5] and mlp["layers"][1]["kernel"]["shape"] == [2, 128]
    assert all(layer[k]["requires_grad"] for layer in mlp["layers"] for k in ("kernel", "bias"))
    back = S.mlp_from_serde(S.from_cbor(S.to_cbor(mlp)))
    assert np.array_equal(back["params"], params) and (back["in_dim"], back["hidden_sizes"], back["out_dim"]) == (5, [128], 2)
    for name in ("actor.cbor", "actor.json"):
        path = str(tmp_path / name)
        actor = S.save_actor(path, params)
        assert list(actor) == ["observation_space", "action_space", "policy_module"] and actor["action_space"] == {}
        obs = actor["observation_space"]["inner"]
        assert list(obs) == ["inner", "remaining"] and obs["remaining"] == {"low": 0.0, "high": 1.0}
        assert list(obs["inner"]) == ["cart_position", "cart_velocity", "pole_angle", "pole_angular_velocity"]
        assert obs["inner"]["cart_position"] == {"low": -2.4, "high": 2.4}
        got = S.load_actor(path)
        assert np.array_equal(got["params"], params) and got["out_dim"] == 2 and got["activation"] == "Relu"
    # unbounded velocity intervals travel as CBOR infinities
    raw = open(str(tmp_path / "actor.cbor"), "rb").read()
    assert S.from_cbor(raw)["observation_space"]["inner"]["inner"]["cart_velocity"] == {"low": -np.inf, "high": np.inf}


def test_tensor_from_def_rejects_bad_input():
    d = S.tensor_def(np.arange(6, dtype=np.float32).reshape(2, 3))
    bad = dict(d, byte_order="BigEndian")
    with pytest.raises(ValueError, match="non-native byte order"):
        S.tensor_from_def(bad)
    with pytest.raises(ValueError):
        S.tensor_from_def(dict(d, shape=[7]))
    with pytest.raises(ValueError):
        S.tensor_from_def({k: v for k, v in d.items() if k != "data"})
    with pytest.raises(ValueError):
        S.mlp_from_serde({"layers": [{"kernel": d, "bias": None}], "activation": "Relu", "output_activation": "Identity"})
