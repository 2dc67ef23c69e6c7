"""Dependency-free reader for frozen TensorFlow GraphDef (.pb) files.

The reference ships its models as frozen GraphDefs
(models-release/master-ckpt-frozen/predict_graph-*.pb,
models-release/supres-40k-swir/superresolve_graph.pb; loaded at
src/download_and_predict_job.py:1785-1826 with tf.import_graph_def).
TensorFlow is not required here: this walks the protobuf wire format directly.

GraphDef{1: NodeDef*}; NodeDef{1:name,2:op,3:input*,5:attr map<string,AttrValue>};
AttrValue{1:list,2:s,3:i,4:f,5:b,6:type,7:shape,8:tensor};
TensorProto{1:dtype,2:shape,4:tensor_content,5:float_val,7:int_val,10:int64_val,...};
TensorShapeProto{2: Dim{1:size}}.
"""
import struct
import numpy as np

_DT = {1: np.float32, 2: np.float64, 3: np.int32, 9: np.int64, 10: np.bool_}


def _varint(b, i):
    r = 0
    s = 0
    while True:
        c = b[i]
        i += 1
        r |= (c & 0x7F) << s
        if c < 0x80:
            return r, i
        s += 7


def _fields(b):
    """Yield (field_number, wire_type, value) over a serialized message."""
    i, n = 0, len(b)
    while i < n:
        key, i = _varint(b, i)
        f, w = key >> 3, key & 7
        if w == 0:
            v, i = _varint(b, i)
        elif w == 1:
            v = b[i:i + 8]; i += 8
        elif w == 2:
            l, i = _varint(b, i)
            v = b[i:i + l]; i += l
        elif w == 5:
            v = b[i:i + 4]; i += 4
        else:
            raise ValueError("unsupported wire type %d" % w)
        yield f, w, v


def _signed(v):
    return v - (1 << 64) if v >= (1 << 63) else v


def _shape(b):
    dims = []
    for f, w, v in _fields(b):
        if f == 2:
            size = 0
            for f2, w2, v2 in _fields(v):
                if f2 == 1:
                    size = _signed(v2)
            dims.append(size)
    return dims


def _packed_varints(v):
    out = []
    i = 0
    while i < len(v):
        x, i = _varint(v, i)
        out.append(_signed(x))
    return out


def _tensor(b):
    dtype, shape, content = 1, [], None
    fvals, ivals, i64vals, bvals = [], [], [], []
    for f, w, v in _fields(b):
        if f == 1:
            dtype = v
        elif f == 2:
            shape = _shape(v)
        elif f == 4:
            content = bytes(v)
        elif f == 5:
            if w == 2:
                fvals.extend(struct.unpack("<%df" % (len(v) // 4), v))
            else:
                fvals.append(struct.unpack("<f", v)[0])
        elif f == 7:
            ivals.extend(_packed_varints(v) if w == 2 else [_signed(v)])
        elif f == 10:
            i64vals.extend(_packed_varints(v) if w == 2 else [_signed(v)])
        elif f == 11:
            bvals.extend(_packed_varints(v) if w == 2 else [v])
    np_dt = _DT.get(dtype)
    if np_dt is None:
        return None
    n = int(np.prod(shape)) if shape else 1
    if content is not None:
        arr = np.frombuffer(content, dtype=np_dt).copy()
    else:
        vals = {np.float32: fvals, np.int32: ivals, np.int64: i64vals,
                np.bool_: bvals, np.float64: fvals}[np_dt]
        if len(vals) == 0:
            arr = np.zeros(n, np_dt)
        elif len(vals) == 1:
            arr = np.full(n, vals[0], np_dt)
        else:
            arr = np.asarray(vals, np_dt)
    return arr.reshape(shape)


def _attr(b):
    """Decode one AttrValue into a python value."""
    for f, w, v in _fields(b):
        if f == 2:
            return bytes(v)
        if f == 3:
            return _signed(v)
        if f == 4:
            return struct.unpack("<f", v)[0]
        if f == 5:
            return bool(v)
        if f == 6:
            return ("dtype", v)
        if f == 7:
            return ("shape", _shape(v))
        if f == 8:
            return _tensor(v)
        if f == 1:  # ListValue: {2:s,3:i(packed),4:f,5:b,6:type}
            out = []
            for f2, w2, v2 in _fields(v):
                if f2 == 3:
                    out.extend(_packed_varints(v2) if w2 == 2 else [_signed(v2)])
                elif f2 == 2:
                    out.append(bytes(v2))
                elif f2 == 4:
                    out.append(struct.unpack("<f", v2)[0])
                elif f2 == 6:
                    out.extend(_packed_varints(v2) if w2 == 2 else [v2])
            return out
    return None


def read_graph(path):
    """Return list of node dicts {name, op, inputs, attr} in file order."""
    with open(path, "rb") as fh:
        buf = memoryview(fh.read())
    nodes = []
    for f, w, v in _fields(buf):
        if f != 1:
            continue
        node = {"name": "", "op": "", "inputs": [], "attr": {}}
        for f2, w2, v2 in _fields(v):
            if f2 == 1:
                node["name"] = bytes(v2).decode()
            elif f2 == 2:
                node["op"] = bytes(v2).decode()
            elif f2 == 3:
                node["inputs"].append(bytes(v2).decode())
            elif f2 == 5:
                k, val = None, None
                for f3, w3, v3 in _fields(v2):
                    if f3 == 1:
                        k = bytes(v3).decode()
                    elif f3 == 2:
                        val = _attr(v3)
                node["attr"][k] = val
        nodes.append(node)
    return nodes


def read_consts(path, float_only=True):
    """{name: ndarray} for every Const node (float32 only by default)."""
    out = {}
    for n in read_graph(path):
        if n["op"] == "Const":
            t = n["attr"].get("value")
            if isinstance(t, np.ndarray) and (not float_only or t.dtype == np.float32):
                out[n["name"]] = t
    return out


if __name__ == "__main__":
    import sys
    from collections import Counter
    nodes = read_graph(sys.argv[1])
    print(len(nodes), "nodes")
    print(Counter(n["op"] for n in nodes).most_common())
    if len(sys.argv) > 2:
        for n in nodes:
            a = {k: (v if not isinstance(v, np.ndarray) else "T%s%s" % (v.dtype, list(v.shape)))
                 for k, v in n["attr"].items() if k not in ("_output_shapes",)}
            print(n["name"], "|", n["op"], "|", n["inputs"], "|", a)
