"""AZ checkpoint container: the reference's `AZ.Save` / `AZ.Load` file (agogo.go:175-209) as a Go `encoding/gob` stream.

What the reference writes (all in-tree, so this part is certain):

    file   = gob.NewEncoder(f).Encode(a.A.NN)            // *dual.Dual implements GobEncoder (dual.go:180-192)
    Dual   -> GobEncode(): a second gob stream, one `enc.Encode(&v)` per Model() node, v a gorgonia.Value interface
    Model(): filters/weights first (conv filters, then the linear weights), then the biases (dual.go:134-141) — the
             engine's az_net_param_desc order.

What lives in un-vendored modules and is restated here FROM MEMORY — **tensor.Dense layout unverified**:

    *tensor.Dense implements GobEncoder (gorgonia.org/tensor v0.9.18, dense_io.go): a third gob stream holding
        Encode(t.Shape())  Encode(t.Strides())  Encode(t.AP.o)  Encode(t.AP.Δ)  Encode(t.mask)  Encode(&data)
    with Shape = named []int, Strides = []int, o / Δ = named uint8 (DataOrder / Triangle), mask = []bool and
    data = interface{} holding []float32; the registered interface names are DENSE_IFACE_NAME / F32S_IFACE_NAME below.

The wire format itself follows the `encoding/gob` package documentation: messages = (uint byte count, payload); payload
= type id (int; negative for a type definition followed by a wireType struct) then the value; unsigned integers are one
byte below 128, else a negated byte count followed by big-endian bytes; signed integers put the sign in bit 0; floats
are byte-reversed float64 bits sent as a uint; structs are (field delta, value)* terminated by 0; a non-struct top-level
value is preceded by a 0 byte; interface values are (concrete type name, type id, byte count, value); user type ids
start at 65.  Type ids are process-global in Go (assigned on first use), so even a correct writer is byte-identical to
a Go-produced file only up to those ids — what must hold is that Go's decoder accepts the stream, which cannot be
checked here (no Go toolchain).  The reader accepts whatever ids the stream defines.

Round trip (writer -> reader) and the framing rules are tested in tests/test_gob_checkpoint.py; `host.AZ.Save/Load` use
this container when the file name does not end in .npz."""
import struct

import numpy as np

# registered names of the concrete types travelling inside interfaces (gob.Register): unverified choices, isolated here
DENSE_IFACE_NAME = "*tensor.Dense"
F32S_IFACE_NAME = "[]float32"

T_BOOL, T_INT, T_UINT, T_FLOAT, T_BYTES, T_STRING, T_COMPLEX, T_INTERFACE = 1, 2, 3, 4, 5, 6, 7, 8
FIRST_USER_ID = 65
# wireType field numbers (encoding/gob/type.go)
WT_ARRAY, WT_SLICE, WT_STRUCT, WT_MAP, WT_GOBENC, WT_BINMARSH, WT_TEXTMARSH = range(7)


# ---------------------------------------------------------------------------------------------------------------------
# primitives
def enc_uint(x):
    if x < 0:
        raise ValueError("uint")
    if x < 128:
        return bytes([x])
    b = x.to_bytes((x.bit_length() + 7) // 8, "big")
    return bytes([256 - len(b)]) + b


def enc_int(i):
    u = (~i << 1) | 1 if i < 0 else i << 1
    return enc_uint(u & ((1 << 64) - 1))


def enc_float(f):
    bits = struct.unpack("<Q", struct.pack("<d", float(f)))[0]
    rev = int.from_bytes(bits.to_bytes(8, "big")[::-1], "big")  # byte-reversed: exponent-first floats become short
    return enc_uint(rev)


def enc_string(s):
    b = s.encode() if isinstance(s, str) else bytes(s)
    return enc_uint(len(b)) + b


class Reader:
    def __init__(self, data):
        self.b, self.i = memoryview(bytes(data)), 0

    def eof(self):
        return self.i >= len(self.b)

    def take(self, n):
        if self.i + n > len(self.b):
            raise ValueError("gob: truncated stream")
        out = bytes(self.b[self.i:self.i + n])
        self.i += n
        return out

    def uint(self):
        c = self.take(1)[0]
        if c < 128:
            return c
        n = 256 - c
        if n > 8:
            raise ValueError("gob: bad uint")
        return int.from_bytes(self.take(n), "big")

    def int(self):
        u = self.uint()
        return ~(u >> 1) if u & 1 else u >> 1

    def float(self):
        rev = self.uint()
        bits = int.from_bytes(rev.to_bytes(8, "big")[::-1], "big")
        return struct.unpack("<d", struct.pack("<Q", bits))[0]

    def string(self):
        return self.take(self.uint())


# ---------------------------------------------------------------------------------------------------------------------
# encoder: one gob stream (type definitions sent once, before first use)
class Encoder:
    def __init__(self):
        self.out = bytearray()
        self.next_id = FIRST_USER_ID
        self.ids = {}

    def _message(self, payload):
        self.out += enc_uint(len(payload)) + payload

    def _common(self, name, tid):  # CommonType{Name, Id}
        return enc_uint(1) + enc_string(name) + enc_uint(1) + enc_int(tid) + enc_uint(0)

    def _define(self, key, name, field, extra=b""):
        """Send wireType{<field>: &T{CommonType{name, id}, extra...}} once; returns the id."""
        if key in self.ids:
            return self.ids[key]
        tid = self.next_id
        self.next_id += 1
        self.ids[key] = tid
        inner = enc_uint(1) + self._common(name, tid) + extra + enc_uint(0)      # T: field 0 = CommonType, then extras
        wire = enc_uint(field + 1) + inner + enc_uint(0)                         # wireType: delta from -1 to `field`
        self._message(enc_int(-tid) + wire)
        return tid

    def gob_encoder_type(self, name):
        return self._define(("gobenc", name), name, WT_GOBENC)

    def slice_type(self, name, elem):
        return self._define(("slice", name), name, WT_SLICE, enc_uint(1) + enc_int(elem))  # sliceType.Elem = field 1

    # ---- values (each = one message)
    def value_gob_encoder(self, name, payload):
        tid = self.gob_encoder_type(name)
        self._message(enc_int(tid) + enc_uint(0) + enc_string(payload))          # singleton: 0 delta, then the bytes

    def value_slice(self, name, elem, items, enc_item):
        tid = self.slice_type(name, elem)
        body = enc_uint(len(items)) + b"".join(enc_item(x) for x in items)
        self._message(enc_int(tid) + enc_uint(0) + body)

    def value_uint(self, x):
        self._message(enc_int(T_UINT) + enc_uint(0) + enc_uint(int(x)))

    def value_interface(self, iface_name, concrete_id, concrete_value, singleton=True):
        """Top-level value of interface type: (name, concrete type id, byte count, [0] value)."""
        val = (enc_uint(0) if singleton else b"") + concrete_value
        body = enc_string(iface_name) + enc_int(concrete_id) + enc_uint(len(val)) + val
        self._message(enc_int(T_INTERFACE) + enc_uint(0) + body)

    def bytes(self):
        return bytes(self.out)


def dense_gob(arr):
    """tensor.Dense.GobEncode of a C-contiguous float32 array (layout restated from memory: UNVERIFIED)."""
    a = np.ascontiguousarray(arr, np.float32)
    shape = list(a.shape) if a.ndim else []
    strides = [int(s // 4) for s in a.strides] if a.ndim else []
    e = Encoder()
    e.value_slice("Shape", T_INT, shape, enc_int)
    e.value_slice("[]int", T_INT, strides, enc_int)
    e.value_uint(0)                                   # AP.o  (DataOrder: row-major, contiguous)
    e.value_uint(0)                                   # AP.Δ  (Triangle: NotTriangle)
    e.value_slice("[]bool", T_BOOL, [], lambda x: enc_uint(1 if x else 0))   # mask
    fid = e.slice_type(F32S_IFACE_NAME, T_FLOAT)
    flat = a.reshape(-1)
    body = enc_uint(flat.size) + enc_floats(flat)
    e.value_interface(F32S_IFACE_NAME, fid, body)
    return e.bytes()


def enc_floats(flat):
    """enc_float over an array, vectorised: per element the float64 bytes low-to-high with the leading zero bytes
    dropped, prefixed by the negated byte count unless the value fits one byte below 128."""
    d = np.ascontiguousarray(flat, "<f8").view(np.uint8).reshape(-1, 8)
    nz = d != 0
    first = np.where(nz.any(axis=1), nz.argmax(axis=1), 7)
    single = (first == 7) & (d[:, 7] < 128)
    m = np.empty((d.shape[0], 9), np.uint8)
    m[:, 0] = (256 - (8 - first)).astype(np.uint8)
    m[:, 1:] = d
    keep = np.empty((d.shape[0], 9), bool)
    keep[:, 0] = ~single
    keep[:, 1:] = np.arange(8)[None, :] >= first[:, None]
    return m[keep].tobytes()


def dual_gob(tensors):
    """dual.Dual.GobEncode (dual.go:180-192): one interface-typed value per Model() tensor."""
    e = Encoder()
    for t in tensors:
        did = e.gob_encoder_type("Dense")
        e.value_interface(DENSE_IFACE_NAME, did, enc_string(dense_gob(t)))
    return e.bytes()


def save_stream(tensors):
    """AZ.Save (agogo.go:175-185): gob.NewEncoder(f).Encode(a.A.NN)."""
    e = Encoder()
    e.value_gob_encoder("Dual", dual_gob(tensors))
    return e.bytes()


# ---------------------------------------------------------------------------------------------------------------------
# decoder (accepts any type ids the stream defines)
class Decoder:
    def __init__(self, data):
        self.r = Reader(data)
        self.types = {}  # id -> ("gobenc", name) | ("slice", name, elem)

    def _parse_typedef(self, r, tid):
        field = -1
        kind = None
        name, elem = "", None
        while True:
            d = r.uint()
            if d == 0:
                break
            field += d
            # nested struct T { CommonType; [Elem] }
            f2 = -1
            while True:
                d2 = r.uint()
                if d2 == 0:
                    break
                f2 += d2
                if f2 == 0:  # CommonType
                    f3 = -1
                    while True:
                        d3 = r.uint()
                        if d3 == 0:
                            break
                        f3 += d3
                        if f3 == 0:
                            name = r.string().decode()
                        elif f3 == 1:
                            r.int()
                        else:
                            raise ValueError("gob: CommonType field %d" % f3)
                elif f2 == 1:
                    elem = r.int()
                else:
                    raise ValueError("gob: unsupported type field %d" % f2)
            kind = field
        if kind == WT_GOBENC:
            self.types[tid] = ("gobenc", name)
        elif kind == WT_SLICE:
            self.types[tid] = ("slice", name, elem)
        else:
            raise ValueError("gob: unsupported wire type %r" % kind)

    def next_value(self):
        """Returns (type id, Reader over the value payload after the id) of the next value message."""
        while True:
            if self.r.eof():
                return None
            n = self.r.uint()
            msg = Reader(self.r.take(n))
            tid = msg.int()
            if tid < 0:
                self._parse_typedef(msg, -tid)
                continue
            return tid, msg

    def _basic(self, r, elem):
        if elem == T_INT:
            return r.int()
        if elem == T_UINT:
            return r.uint()
        if elem == T_BOOL:
            return bool(r.uint())
        if elem == T_FLOAT:
            return r.float()
        raise ValueError("gob: unsupported element type %d" % elem)

    def read_slice(self, r, tid):
        kind = self.types.get(tid)
        if not kind or kind[0] != "slice":
            raise ValueError("gob: type %d is not a slice" % tid)
        n = r.uint()
        if kind[2] == T_FLOAT:  # the payload of a network
            buf, i, out = r.b, r.i, np.empty(n, np.float64)
            raw = bytearray(8)
            for k in range(n):
                c = buf[i]
                if c < 128:
                    raw[:] = b"\0" * 7 + bytes([c]); i += 1
                else:
                    ln = 256 - c
                    raw[:] = b"\0" * (8 - ln) + bytes(buf[i + 1:i + 1 + ln]); i += 1 + ln
                out[k] = struct.unpack("<d", raw)[0]
            r.i = i
            return out
        return [self._basic(r, kind[2]) for _ in range(n)]

    def read_singleton_header(self, r):
        if r.uint() != 0:
            raise ValueError("gob: expected the singleton's zero delta")

    def read_interface(self, r):
        """-> (registered name, concrete type id, Reader over the concrete value incl. its singleton delta)"""
        name = r.string().decode()
        tid = r.int()
        n = r.uint()
        return name, tid, Reader(r.take(n))


def parse_dense(payload):
    d = Decoder(payload)
    tid, r = d.next_value(); d.read_singleton_header(r); shape = d.read_slice(r, tid)
    tid, r = d.next_value(); d.read_singleton_header(r); strides = d.read_slice(r, tid)
    for _ in range(2):  # o, Δ
        tid, r = d.next_value(); d.read_singleton_header(r); r.uint()
    tid, r = d.next_value(); d.read_singleton_header(r); d.read_slice(r, tid)  # mask
    tid, r = d.next_value()
    if tid != T_INTERFACE:
        raise ValueError("gob: Dense data is not an interface value")
    d.read_singleton_header(r)
    _, cid, vr = d.read_interface(r)
    d.read_singleton_header(vr)
    data = np.array(d.read_slice(vr, cid), np.float32)
    want = int(np.prod(shape)) if shape else 1
    if data.size != want:
        raise ValueError("gob: Dense data length %d does not match shape %r" % (data.size, shape))
    expect = [int(np.prod(shape[i + 1:])) for i in range(len(shape))]
    if list(strides) != expect:
        raise ValueError("gob: non-contiguous Dense (strides %r)" % (strides,))
    return data.reshape(shape) if shape else data.reshape(())


def load_stream(data):
    """AZ.Load's decode of one net: -> list of float32 arrays in Model() order."""
    d = Decoder(data)
    got = d.next_value()
    if got is None:
        raise ValueError("gob: empty checkpoint")
    tid, r = got
    if d.types.get(tid, ("",))[0] != "gobenc":
        raise ValueError("gob: the checkpoint's top-level value is not a GobEncoder type")
    d.read_singleton_header(r)
    inner = Decoder(r.string())
    out = []
    while True:
        got = inner.next_value()
        if got is None:
            break
        tid, r = got
        if tid != T_INTERFACE:
            raise ValueError("gob: Model() entry is not an interface value")
        inner.read_singleton_header(r)
        _, cid, vr = inner.read_interface(r)
        if inner.types.get(cid, ("",))[0] != "gobenc":
            raise ValueError("gob: Model() entry is not a GobEncoder (tensor.Dense)")
        inner.read_singleton_header(vr)
        out.append(parse_dense(vr.string()))
    return out
