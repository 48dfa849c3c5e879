"""Actor / module wire format of the reference: serde `TensorDef` inside CBOR or JSON.

The reference saves a trained actor with `serde_cbor::to_writer(file, &agent.actor(ActorMode::Evaluation))`
(examples/cartpole-trpo.rs:69-75) and loads it back for evaluation (:79-91).  What is written is the serde data
model of

    PolicyActor { observation_space: NonEmptyFeatures { inner: <obs space> }, action_space, policy_module }
                                                              (src/torch/agents/policies/actor.rs:9-14)
    Mlp { layers: Vec<Linear>, activation, output_activation }                (torch/modules/ff/mlp.rs:45-50)
    Linear { kernel: TensorDef, bias: Option<TensorDef> }                     (torch/modules/ff/linear.rs:43-50)
    TensorDef { kind, shape, requires_grad, byte_order, data: bytes }         (torch/serialize.rs:66-79)

This module restates that data model on plain Python values (dict = struct with named fields in declaration order,
list = seq, str = unit enum variant, bytes = byte string, None = Option::None) and the two self-describing formats
the reference uses: CBOR (RFC 8949, as serde_cbor 0.11 writes it: structs as maps keyed by field name, unit variants
as text, definite lengths) and JSON (as serde_json writes it: byte strings as arrays of numbers).

Pinned by the reference's own known-answer tests: the four serde_test token streams of torch/serialize.rs:176-340
(`tests/test_serialize.py` transcribes them).  serde_cbor and serde_json themselves are third-party crates that are
not vendored under the reference; their encodings are restated from RFC 8949 / RFC 8259 and the decoders here accept
every float / integer width, so a file written by the Rust side loads whichever width it chose.
"""
from __future__ import annotations

import json
import struct
from collections import OrderedDict

import numpy as np

# torch/serialize.rs:14-31 (KindDef variant names) <-> numpy dtypes this path uses
_KIND_OF_DTYPE = {
    np.dtype(np.uint8): "Uint8", np.dtype(np.int8): "Int8", np.dtype(np.int16): "Int16", np.dtype(np.int32): "Int",
    np.dtype(np.int64): "Int64", np.dtype(np.float16): "Half", np.dtype(np.float32): "Float",
    np.dtype(np.float64): "Double", np.dtype(np.bool_): "Bool",
}
_DTYPE_OF_KIND = {v: k for k, v in _KIND_OF_DTYPE.items()}
NATIVE_BYTE_ORDER = "LittleEndian" if np.little_endian else "BigEndian"  # serialize.rs:50-58


# ------------------------------------------------------------------------------------------------
# TensorDef (torch/serialize.rs:66-115)
# ------------------------------------------------------------------------------------------------
def tensor_def(array, requires_grad: bool = False) -> OrderedDict:
    """`TensorDef::from(&Tensor)`: kind, shape, requires_grad, native byte order, a copy of the data bytes."""
    a = np.asarray(array)
    if a.dtype not in _KIND_OF_DTYPE:
        raise TypeError(f"no tch Kind for dtype {a.dtype}")
    # (np.ascontiguousarray would turn a 0-d tensor into shape [1]; tobytes() already emits C order)
    return OrderedDict(kind=_KIND_OF_DTYPE[a.dtype], shape=[int(n) for n in a.shape], requires_grad=bool(requires_grad),
                       byte_order=NATIVE_BYTE_ORDER, data=a.tobytes(order="C"))


def tensor_from_def(d) -> tuple[np.ndarray, bool]:
    """`Tensor::from(&TensorDef)`: refuses a foreign byte order exactly like the reference's assert (serialize.rs:107-111)."""
    for field in ("kind", "shape", "requires_grad", "byte_order", "data"):
        if field not in d:
            raise ValueError(f"TensorDef: missing field `{field}`")
    if d["byte_order"] != NATIVE_BYTE_ORDER:
        raise ValueError("data has non-native byte order")
    if d["kind"] not in _DTYPE_OF_KIND:
        raise ValueError(f"TensorDef: unsupported kind {d['kind']!r}")
    data = bytes(d["data"]) if not isinstance(d["data"], (bytes, bytearray)) else d["data"]
    shape = tuple(int(n) for n in d["shape"])
    dtype = _DTYPE_OF_KIND[d["kind"]]
    if int(np.prod(shape, dtype=np.int64)) * dtype.itemsize != len(data):
        raise ValueError("TensorDef: data length does not match kind and shape")
    return np.frombuffer(data, dtype=dtype).reshape(shape).copy(), bool(d["requires_grad"])


def tensor_def_tokens(d) -> list:
    """The serde_test token stream of a TensorDef (what torch/serialize.rs:176-340 asserts)."""
    toks = [("Struct", "TensorDef", 5), ("Str", "kind"), ("UnitVariant", "KindDef", d["kind"]), ("Str", "shape"),
            ("Seq", len(d["shape"]))]
    toks += [("I64", int(n)) for n in d["shape"]]
    toks += [("SeqEnd",), ("Str", "requires_grad"), ("Bool", bool(d["requires_grad"])), ("Str", "byte_order"),
             ("UnitVariant", "ByteOrder", d["byte_order"]), ("Str", "data"), ("BorrowedBytes", bytes(d["data"])), ("StructEnd",)]
    return toks


# ------------------------------------------------------------------------------------------------
# CBOR (RFC 8949) for the serde data model
# ------------------------------------------------------------------------------------------------
def _head(major: int, n: int) -> bytes:
    if n < 24:
        return bytes([major << 5 | n])
    for code, fmt, lim in ((24, ">B", 1 << 8), (25, ">H", 1 << 16), (26, ">I", 1 << 32), (27, ">Q", 1 << 64)):
        if n < lim:
            return bytes([major << 5 | code]) + struct.pack(fmt, n)
    raise OverflowError("integer too large for CBOR")


def _float_bytes(x: float) -> bytes:
    # shortest IEEE width that holds the value exactly (RFC 8949 4.2.2 preferred serialization; serde_cbor packs likewise)
    if x != x:
        return b"\xf9\x7e\x00"
    for code, fmt in ((0xF9, ">e"), (0xFA, ">f")):
        try:
            packed = struct.pack(fmt, x)
        except (OverflowError, struct.error):
            continue
        if struct.unpack(fmt, packed)[0] == x:
            return bytes([code]) + packed
    return b"\xfb" + struct.pack(">d", x)


def to_cbor(v) -> bytes:
    if v is None:
        return b"\xf6"
    if v is True:
        return b"\xf5"
    if v is False:
        return b"\xf4"
    if isinstance(v, (int, np.integer)):
        v = int(v)
        return _head(0, v) if v >= 0 else _head(1, -1 - v)
    if isinstance(v, (float, np.floating)):
        return _float_bytes(float(v))
    if isinstance(v, (bytes, bytearray)):
        return _head(2, len(v)) + bytes(v)
    if isinstance(v, str):
        b = v.encode("utf-8")
        return _head(3, len(b)) + b
    if isinstance(v, (list, tuple)):
        return _head(4, len(v)) + b"".join(to_cbor(x) for x in v)
    if isinstance(v, dict):
        return _head(5, len(v)) + b"".join(to_cbor(k) + to_cbor(x) for k, x in v.items())
    raise TypeError(f"cannot encode {type(v)} as CBOR")


class _Reader:
    def __init__(self, data: bytes):
        self.d, self.i = memoryview(data), 0

    def take(self, n: int) -> bytes:
        if self.i + n > len(self.d):
            raise ValueError("CBOR: unexpected end of input")
        out = bytes(self.d[self.i:self.i + n])
        self.i += n
        return out

    def arg(self, info: int):
        if info < 24:
            return info
        if info in (24, 25, 26, 27):
            return int.from_bytes(self.take(1 << (info - 24)), "big")
        if info == 31:
            return None  # indefinite length
        raise ValueError("CBOR: reserved additional information")

    def value(self):
        b = self.take(1)[0]
        major, info = b >> 5, b & 31
        if major == 7:
            if info == 20:
                return False
            if info == 21:
                return True
            if info in (22, 23):
                return None
            if info == 25:
                return struct.unpack(">e", self.take(2))[0]
            if info == 26:
                return struct.unpack(">f", self.take(4))[0]
            if info == 27:
                return struct.unpack(">d", self.take(8))[0]
            raise ValueError(f"CBOR: unsupported simple value {info}")
        n = self.arg(info)
        if major == 0:
            return n
        if major == 1:
            return -1 - n
        if major in (2, 3):
            if n is None:  # indefinite: concatenated definite chunks up to the break byte
                chunks = []
                while self.d[self.i] != 0xFF:
                    chunks.append(self.value())
                self.i += 1
                return b"".join(chunks) if major == 2 else "".join(chunks)
            raw = self.take(n)
            return raw if major == 2 else raw.decode("utf-8")
        if major == 4:
            out = []
            if n is None:
                while self.d[self.i] != 0xFF:
                    out.append(self.value())
                self.i += 1
            else:
                out = [self.value() for _ in range(n)]
            return out
        if major == 5:
            out = OrderedDict()
            if n is None:
                while self.d[self.i] != 0xFF:
                    k = self.value()
                    out[k] = self.value()
                self.i += 1
            else:
                for _ in range(n):
                    k = self.value()
                    out[k] = self.value()
            return out
        if major == 6:  # tag: keep the tagged value
            return self.value()
        raise ValueError("CBOR: bad major type")


def from_cbor(data: bytes):
    r = _Reader(data)
    v = r.value()
    if r.i != len(data):
        raise ValueError("CBOR: trailing bytes")
    return v


# ------------------------------------------------------------------------------------------------
# JSON as serde_json writes the same model (bytes -> array of numbers, non-finite floats -> null)
# ------------------------------------------------------------------------------------------------
def _jsonable(v):
    if isinstance(v, (bytes, bytearray)):
        return list(v)
    if isinstance(v, dict):
        return OrderedDict((k, _jsonable(x)) for k, x in v.items())
    if isinstance(v, (list, tuple)):
        return [_jsonable(x) for x in v]
    if isinstance(v, (float, np.floating)) and not np.isfinite(v):
        return None
    if isinstance(v, np.integer):
        return int(v)
    if isinstance(v, np.floating):
        return float(v)
    return v


def to_json(v) -> str:
    return json.dumps(_jsonable(v), separators=(",", ":"))


def from_json(text: str):
    return json.loads(text, object_pairs_hook=OrderedDict)


# ------------------------------------------------------------------------------------------------
# Modules and the policy actor
# ------------------------------------------------------------------------------------------------
_ACTIVATIONS = ("Identity", "Relu", "Sigmoid", "Tanh")  # torch/modules/ff/activation.rs:11-20 == RL_ACT_* order


def mlp_to_serde(flat_params, in_dim: int, hidden_sizes, out_dim: int, activation: str = "Relu",
                 output_activation: str = "Identity") -> OrderedDict:
    """`Mlp` (mlp.rs:45-50) from the flat parameter vector in `Module::variables()` order (kernel [out, in] row-major,
    bias [out], layer by layer -- linear.rs:108-110), trainable tensors flagged `requires_grad`."""
    flat = np.asarray(flat_params, np.float32)
    dims = [in_dim] + list(hidden_sizes) + [out_dim]
    layers, at = [], 0
    for i, o in zip(dims[:-1], dims[1:]):
        kernel = flat[at:at + o * i].reshape(o, i)
        at += o * i
        bias = flat[at:at + o]
        at += o
        layers.append(OrderedDict(kernel=tensor_def(kernel, True), bias=tensor_def(bias, True)))
    if at != flat.size:
        raise ValueError(f"expected {at} parameters for dims {dims}, got {flat.size}")
    if activation not in _ACTIVATIONS or output_activation not in _ACTIVATIONS:
        raise ValueError("unknown activation")
    return OrderedDict(layers=layers, activation=activation, output_activation=output_activation)


def mlp_from_serde(d) -> dict:
    """Inverse of `mlp_to_serde`: flat f32 parameters + the layer sizes; a layer without bias is refused (the kernels
    of this library are built for LinearConfig::default(), bias_init = Some)."""
    flat, dims = [], []
    for layer in d["layers"]:
        kernel, _ = tensor_from_def(layer["kernel"])
        if layer.get("bias") is None:
            raise ValueError("Linear without bias is not supported by relearn_b200 modules")
        bias, _ = tensor_from_def(layer["bias"])
        if kernel.ndim != 2 or bias.shape != (kernel.shape[0],) or kernel.dtype != np.float32 or bias.dtype != np.float32:
            raise ValueError("Linear: kernel must be f32 [out, in] and bias f32 [out]")
        if dims and dims[-1] != kernel.shape[1]:
            raise ValueError("Mlp: consecutive layer shapes do not chain")
        if not dims:
            dims.append(int(kernel.shape[1]))
        dims.append(int(kernel.shape[0]))
        flat += [kernel.reshape(-1), bias]
    return {"params": np.concatenate(flat).astype(np.float32), "in_dim": dims[0], "hidden_sizes": dims[1:-1], "out_dim": dims[-1],
            "activation": d["activation"], "output_activation": d["output_activation"]}


def _interval(low: float, high: float) -> OrderedDict:
    return OrderedDict(low=float(low), high=float(high))  # spaces/interval.rs:14-18


def cartpole_observation_space(max_pos: float = 2.4, max_angle: float = float(np.radians(12.0)), step_limit: bool = True):
    """`NonEmptyFeatures<StepLimitObsSpace<CartPolePhysicalStateSpace>>` (nonempty_features.rs:22-24, step_limit.rs:133-138,
    cartpole.rs:73-82,273-284): the observation space stored next to a CartPole policy."""
    phys = OrderedDict(cart_position=_interval(-max_pos, max_pos), cart_velocity=_interval(-np.inf, np.inf),
                       pole_angle=_interval(-max_angle, max_angle), pole_angular_velocity=_interval(-np.inf, np.inf))
    inner = OrderedDict(inner=phys, remaining=_interval(0.0, 1.0)) if step_limit else phys
    return OrderedDict(inner=inner)


def policy_actor_to_serde(observation_space, mlp) -> OrderedDict:
    """`PolicyActor` (policies/actor.rs:9-14).  `IndexedTypeSpace<Push>` has no serialised field (indexed_type.rs:57-64:
    its PhantomData is `#[serde(skip)]`), so the action space is an empty struct."""
    return OrderedDict(observation_space=observation_space, action_space=OrderedDict(), policy_module=mlp)


def save_actor(path: str, flat_params, in_dim: int = 5, hidden_sizes=(128,), out_dim: int = 2, observation_space=None):
    """Write `actor.cbor` (or `.json`) for a CartPole + VisibleStepLimit MLP policy, loadable by
    `cargo run --example cartpole-trpo <path>` (examples/cartpole-trpo.rs:79-91)."""
    actor = policy_actor_to_serde(observation_space or cartpole_observation_space(),
                                  mlp_to_serde(flat_params, in_dim, list(hidden_sizes), out_dim))
    if path.endswith(".json"):
        with open(path, "w") as f:
            f.write(to_json(actor))
    else:
        with open(path, "wb") as f:
            f.write(to_cbor(actor))
    return actor


def load_actor(path: str) -> dict:
    """Read an actor written by the reference (or by `save_actor`): the policy module as flat parameters + sizes and
    the stored spaces, untouched."""
    if path.endswith(".json"):
        with open(path) as f:
            actor = from_json(f.read())
    else:
        with open(path, "rb") as f:
            actor = from_cbor(f.read())
    for field in ("observation_space", "action_space", "policy_module"):
        if field not in actor:
            raise ValueError(f"PolicyActor: missing field `{field}`")
    out = mlp_from_serde(actor["policy_module"])
    out["observation_space"], out["action_space"] = actor["observation_space"], actor["action_space"]
    return out
