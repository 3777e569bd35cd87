"""Checkpoint interchange with the reference trainer -- no TensorFlow needed.

The reference saves with `tf.train.Saver` through `tf.train.Supervisor` (train.py:415-445): TF-1.x "V2" checkpoints,
i.e. a *tensor bundle*

    <prefix>.index                    an SSTable (LevelDB table format, uncompressed blocks): key "" -> BundleHeaderProto,
                                      key <variable name> -> BundleEntryProto {dtype, shape, shard_id, offset, size, crc32c}
    <prefix>.data-00000-of-00001      the raw little-endian tensor bytes at those offsets

and next to them `model_flags.json` (train.py:390-411).  Variable names carry the tower scope of train.py:277
(`tower/input_bn/gamma`, ...), Adam slots are `<var>/Adam` and `<var>/Adam_1`, and `global_step`, `beta1_power`,
`beta2_power` ride along.  This module reads and writes that format from its published layout (TensorFlow
core/util/tensor_bundle + core/lib/io/table*, pinned to the format revision every TF 1.x release wrote) and maps the
names onto `variables.VariableStore` (the state-dict keys ARE the TF names).

PARITY UNPINNED for the file format: no TensorFlow and no reference checkpoint exist in this environment, so the
reader is verified against the writer (round trip), against the format's own invariants (footer magic, block CRCs,
known CRC-32C vectors) and against a byte-level fixture assembled by hand in tests/test_checkpoint_cpu.py.
"""
from __future__ import annotations

import ctypes as C
import json
import os
import struct
from typing import Dict, Iterable, Optional, Tuple

import numpy as np

TABLE_MAGIC = 0xDB4775248B80FB57
RESTART_INTERVAL = 16
BLOCK_SIZE = 4096          # table::Options default used by BundleWriter
MASK_DELTA = 0xA282EAD8

# tensorflow/core/framework/types.proto
_DT = {1: np.float32, 2: np.float64, 3: np.int32, 4: np.uint8, 5: np.int16, 6: np.int8, 9: np.int64, 10: np.bool_,
       17: np.uint16, 19: np.float16, 22: np.uint32, 23: np.uint64}
_DT_OF = {np.dtype(v): k for k, v in _DT.items()}


# ------------------------------------------------------------------------------------------------
# CRC-32C (host utility of the C-ABI library; a pure-Python fallback keeps the reader usable for the small index)
# ------------------------------------------------------------------------------------------------
_crc_fn = None


def _py_crc32c(crc: int, data: bytes) -> int:
    global _PY_TAB
    try:
        tab = _PY_TAB
    except NameError:
        tab = []
        for i in range(256):
            c = i
            for _ in range(8):
                c = (c >> 1) ^ 0x82F63B78 if c & 1 else c >> 1
            tab.append(c)
        _PY_TAB = tab
    c = crc ^ 0xFFFFFFFF
    for b in data:
        c = tab[(c ^ b) & 0xFF] ^ (c >> 8)
    return c ^ 0xFFFFFFFF


def crc32c(data, crc: int = 0) -> int:
    """CRC-32C of a bytes-like / contiguous numpy array, continued from `crc`."""
    global _crc_fn
    if _crc_fn is None:
        try:
            from . import _lib
            lib = _lib.load()
            lib.lpm_crc32c.restype = C.c_uint
            lib.lpm_crc32c.argtypes = [C.c_uint, C.c_void_p, C.c_ulonglong]
            _crc_fn = lib.lpm_crc32c
        except Exception:          # library not built: index blocks are a few KB, tensors are then not verified
            _crc_fn = False
    if isinstance(data, np.ndarray):
        if _crc_fn:
            a = np.ascontiguousarray(data)
            return int(_crc_fn(crc, a.ctypes.data, a.nbytes))
        data = data.tobytes()
    if _crc_fn:
        buf = bytes(data)
        return int(_crc_fn(crc, buf, len(buf)))
    return _py_crc32c(crc, bytes(data))


def have_native_crc() -> bool:
    crc32c(b"")
    return bool(_crc_fn)


def mask_crc(c: int) -> int:
    """crc32c::Mask: rotate right by 15 and add a constant (stored CRCs are masked)."""
    return (((c >> 15) | (c << 17)) + MASK_DELTA) & 0xFFFFFFFF


# ------------------------------------------------------------------------------------------------
# varints / minimal protobuf
# ------------------------------------------------------------------------------------------------
def _put_varint(n: int) -> bytes:
    out = bytearray()
    while n >= 0x80:
        out.append((n & 0x7F) | 0x80)
        n >>= 7
    out.append(n)
    return bytes(out)


def _get_varint(buf, pos: int) -> Tuple[int, int]:
    shift, val = 0, 0
    while True:
        b = buf[pos]
        pos += 1
        val |= (b & 0x7F) << shift
        if b < 0x80:
            return val, pos
        shift += 7
        if shift > 70:
            raise ValueError("malformed varint")


def _pb_fields(buf) -> Iterable[Tuple[int, int, object]]:
    """(field number, wire type, value) of one protobuf message."""
    pos, n = 0, len(buf)
    while pos < n:
        key, pos = _get_varint(buf, pos)
        f, wt = key >> 3, key & 7
        if wt == 0:
            v, pos = _get_varint(buf, pos)
        elif wt == 1:
            v = struct.unpack_from("<Q", buf, pos)[0]; pos += 8
        elif wt == 2:
            ln, pos = _get_varint(buf, pos)
            v = bytes(buf[pos:pos + ln]); pos += ln
        elif wt == 5:
            v = struct.unpack_from("<I", buf, pos)[0]; pos += 4
        else:
            raise ValueError(f"unsupported protobuf wire type {wt}")
        yield f, wt, v


def _pb_key(field: int, wt: int) -> bytes:
    return _put_varint((field << 3) | wt)


def _entry_proto(dtype: int, shape, offset: int, size: int, crc_masked: int) -> bytes:
    """BundleEntryProto: dtype=1, shape=2 (TensorShapeProto.dim=2 {size=1}), shard_id=3, offset=4, size=5, crc32c=6."""
    dims = b"".join(_pb_key(2, 2) + _put_varint(len(d)) + d for d in (_pb_key(1, 0) + _put_varint(int(s)) for s in shape))
    out = _pb_key(1, 0) + _put_varint(dtype) + _pb_key(2, 2) + _put_varint(len(dims)) + dims
    if offset:
        out += _pb_key(4, 0) + _put_varint(offset)
    out += _pb_key(5, 0) + _put_varint(size) + _pb_key(6, 5) + struct.pack("<I", crc_masked)
    return out


def _parse_entry(buf) -> dict:
    e = {"dtype": 0, "shape": [], "shard_id": 0, "offset": 0, "size": 0, "crc32c": 0, "slices": 0}
    for f, wt, v in _pb_fields(buf):
        if f == 1:
            e["dtype"] = v
        elif f == 2:
            for f2, _, v2 in _pb_fields(v):
                if f2 == 2:
                    size = 0
                    for f3, _, v3 in _pb_fields(v2):
                        if f3 == 1:
                            size = v3 - (1 << 64) if v3 >= (1 << 63) else v3
                    e["shape"].append(size)
                elif f2 == 3 and v2:
                    raise ValueError("tensor of unknown rank in checkpoint")
        elif f == 3:
            e["shard_id"] = v
        elif f == 4:
            e["offset"] = v
        elif f == 5:
            e["size"] = v
        elif f == 6:
            e["crc32c"] = v
        elif f == 7:
            e["slices"] += 1
    return e


# ------------------------------------------------------------------------------------------------
# SSTable (LevelDB table format as used by tensorflow/core/lib/io/table*)
# ------------------------------------------------------------------------------------------------
def _read_block(buf, offset: int, size: int, verify: bool) -> bytes:
    data, ctype = buf[offset:offset + size], buf[offset + size]
    stored = struct.unpack_from("<I", buf, offset + size + 1)[0]
    if verify and mask_crc(crc32c(bytes(buf[offset:offset + size + 1]))) != stored:
        raise ValueError(f"checkpoint index: block checksum mismatch at offset {offset}")
    if ctype != 0:
        raise NotImplementedError("compressed (snappy) index blocks: TensorFlow's BundleWriter never writes them")
    return bytes(data)


def _block_entries(block: bytes):
    n_restarts = struct.unpack_from("<I", block, len(block) - 4)[0]
    end = len(block) - 4 - 4 * n_restarts
    pos, key = 0, b""
    while pos < end:
        shared, pos = _get_varint(block, pos)
        non_shared, pos = _get_varint(block, pos)
        vlen, pos = _get_varint(block, pos)
        key = key[:shared] + block[pos:pos + non_shared]
        pos += non_shared
        yield key, block[pos:pos + vlen]
        pos += vlen


def _table_items(buf, verify=True):
    if len(buf) < 48 or struct.unpack_from("<Q", buf, len(buf) - 8)[0] != TABLE_MAGIC:
        raise ValueError("not a TensorFlow checkpoint index (bad table magic)")
    footer = buf[len(buf) - 48:]
    _, p = _get_varint(footer, 0)          # metaindex handle (offset, size): unused
    _, p = _get_varint(footer, p)
    ioff, p = _get_varint(footer, p)
    isize, p = _get_varint(footer, p)
    for _, handle in _block_entries(_read_block(buf, ioff, isize, verify)):
        off, q = _get_varint(handle, 0)
        size, _ = _get_varint(handle, q)
        yield from _block_entries(_read_block(buf, off, size, verify))


class _BlockBuilder:
    def __init__(self):
        self.buf, self.restarts, self.count, self.last = bytearray(), [0], 0, b""

    def add(self, key: bytes, value: bytes):
        shared = 0
        if self.count % RESTART_INTERVAL == 0 and self.count:
            self.restarts.append(len(self.buf))
        elif self.count:
            m = min(len(key), len(self.last))
            while shared < m and key[shared] == self.last[shared]:
                shared += 1
        self.buf += _put_varint(shared) + _put_varint(len(key) - shared) + _put_varint(len(value)) + key[shared:] + value
        self.last, self.count = key, self.count + 1

    def size(self):
        return len(self.buf) + 4 * len(self.restarts) + 4

    def finish(self) -> bytes:
        return bytes(self.buf) + b"".join(struct.pack("<I", r) for r in self.restarts) + struct.pack("<I", len(self.restarts))


def _write_table(path: str, items):
    """items: sorted (key bytes, value bytes).  Uncompressed blocks, restart interval 16, 4 KB target block size."""
    out = bytearray()
    index = _BlockBuilder()

    def emit(block: bytes) -> bytes:
        off = len(out)
        out.extend(block + b"\x00")
        out.extend(struct.pack("<I", mask_crc(crc32c(block + b"\x00"))))
        return _put_varint(off) + _put_varint(len(block))

    cur, last_key = _BlockBuilder(), None
    for key, value in items:
        if last_key is not None and key <= last_key:
            raise ValueError("checkpoint keys must be added in strictly increasing order")
        cur.add(key, value)
        last_key = key
        if cur.size() >= BLOCK_SIZE:
            index.add(last_key, emit(cur.finish()))      # the last key itself is a valid separator
            cur = _BlockBuilder()
    if cur.count:
        index.add(last_key, emit(cur.finish()))
    meta = emit(_BlockBuilder().finish())
    idx = emit(index.finish())
    footer = meta + idx
    out.extend(footer + b"\x00" * (40 - len(footer)) + struct.pack("<Q", TABLE_MAGIC))
    with open(path, "wb") as f:
        f.write(out)


# ------------------------------------------------------------------------------------------------
# tensor bundle
# ------------------------------------------------------------------------------------------------
def list_tf_checkpoint(prefix: str, verify: bool = True) -> Dict[str, dict]:
    """{variable name: {dtype, shape, shard_id, offset, size, crc32c}} of <prefix>.index."""
    with open(prefix + ".index", "rb") as f:
        buf = f.read()
    entries, header_seen = {}, False
    for key, value in _table_items(buf, verify):
        if key == b"":
            header_seen = True
            for f_, _, v in _pb_fields(value):
                if f_ == 2 and v != 0:
                    raise NotImplementedError("big-endian tensor bundle")
            continue
        entries[key.decode("utf-8")] = _parse_entry(value)
    if not header_seen:
        raise ValueError("checkpoint index without a bundle header entry")
    return entries


def read_tf_checkpoint(prefix: str, names: Optional[Iterable[str]] = None, verify: bool = True) -> Dict[str, np.ndarray]:
    """Load tensors of a TF-1.x V2 checkpoint (`prefix` as in the `checkpoint` state file, e.g. .../model.ckpt-1000).
    verify: check index-block checksums and (with the native CRC) every tensor's CRC-32C."""
    entries = list_tf_checkpoint(prefix, verify)
    want = set(names) if names is not None else None
    shards = {e["shard_id"] for e in entries.values()}
    num_shards = max(shards) + 1 if shards else 1
    verify_data = verify and have_native_crc()
    out = {}
    files = {}
    try:
        for name, e in entries.items():
            if want is not None and name not in want:
                continue
            if e["slices"]:
                raise NotImplementedError(f"{name}: partitioned (sliced) variables are not used by the reference trainer")
            if e["dtype"] not in _DT:
                continue                                    # strings etc. (not variables of this model)
            sid = e["shard_id"]
            if sid not in files:
                files[sid] = open(f"{prefix}.data-{sid:05d}-of-{num_shards:05d}", "rb")
            f = files[sid]
            f.seek(e["offset"])
            raw = f.read(e["size"])
            if len(raw) != e["size"]:
                raise ValueError(f"{name}: data file truncated")
            arr = np.frombuffer(raw, dtype=np.dtype(_DT[e["dtype"]]).newbyteorder("<")).reshape(e["shape"])
            if verify_data and mask_crc(crc32c(arr)) != e["crc32c"]:
                raise ValueError(f"{name}: tensor checksum mismatch")
            out[name] = arr
    finally:
        for f in files.values():
            f.close()
    if want is not None and want - set(out):
        raise KeyError(f"not in checkpoint: {sorted(want - set(out))[:5]}")
    return out


def write_tf_checkpoint(prefix: str, tensors: Dict[str, np.ndarray]):
    """Write {name: array} as a single-shard TF-1.x V2 checkpoint that `tf.train.Saver.restore` accepts."""
    os.makedirs(os.path.dirname(os.path.abspath(prefix)), exist_ok=True)
    header = _pb_key(1, 0) + _put_varint(1) + _pb_key(3, 2) + _put_varint(2) + _pb_key(1, 0) + _put_varint(1)
    items = [(b"", header)]                              # num_shards=1, little endian (default), version.producer=1
    offset = 0
    with open(prefix + ".data-00000-of-00001", "wb") as f:
        for name in sorted(tensors, key=lambda s: s.encode("utf-8")):
            a = np.asarray(tensors[name], order="C")        # (ascontiguousarray would promote scalars to shape (1,))
            if a.dtype not in _DT_OF:
                raise TypeError(f"{name}: dtype {a.dtype} has no TensorFlow counterpart here")
            a = a.astype(a.dtype.newbyteorder("<"), copy=False)
            f.write(a.tobytes() if a.ndim == 0 else memoryview(a).cast("B"))
            items.append((name.encode("utf-8"), _entry_proto(_DT_OF[a.dtype], a.shape, offset, a.nbytes, mask_crc(crc32c(a)))))
            offset += a.nbytes
    _write_table(prefix + ".index", items)


def latest_checkpoint(train_dir: str) -> Optional[str]:
    """tf.train.latest_checkpoint: the `model_checkpoint_path` of <train_dir>/checkpoint (text CheckpointState)."""
    path = os.path.join(train_dir, "checkpoint")
    if not os.path.exists(path):
        return None
    for line in open(path):
        if line.startswith("model_checkpoint_path:"):
            p = line.split(":", 1)[1].strip().strip('"')
            return p if os.path.isabs(p) else os.path.join(train_dir, p)
    return None


def update_checkpoint_state(train_dir: str, prefix: str):
    name = os.path.basename(prefix)
    with open(os.path.join(train_dir, "checkpoint"), "w") as f:
        f.write(f'model_checkpoint_path: "{name}"\nall_model_checkpoint_paths: "{name}"\n')


# ------------------------------------------------------------------------------------------------
# model_flags.json (train.py:390-411)
# ------------------------------------------------------------------------------------------------
def model_flags(model="NetVladV1", feature_sizes="1024,128", feature_names="rgb,audio", frame_features=True,
                label_loss="CrossEntropyLoss") -> dict:
    return {"model": model, "feature_sizes": feature_sizes, "feature_names": feature_names,
            "frame_features": frame_features, "label_loss": label_loss}


def write_model_flags(train_dir: str, flags_dict: dict):
    """Same behaviour as train.py:397-411: an existing file with different flags is an error."""
    os.makedirs(train_dir, exist_ok=True)
    path = os.path.join(train_dir, "model_flags.json")
    if os.path.exists(path):
        existing = json.load(open(path))
        if existing != flags_dict:
            raise ValueError(f"Model flags do not match existing file {path}. Ran with {flags_dict}, previously {existing}")
        return
    with open(path, "w") as f:
        f.write(json.dumps(flags_dict))


def read_model_flags(train_dir: str) -> dict:
    return json.load(open(os.path.join(train_dir, "model_flags.json")))


# ------------------------------------------------------------------------------------------------
# VariableStore / Trainer <-> checkpoint
# ------------------------------------------------------------------------------------------------
TOWER = "tower/"           # tf.variable_scope("tower", reuse=...) around create_model (train.py:277)


def load_into_store(store, prefix: str, trainer=None, strict: bool = True, verify: bool = True) -> dict:
    """Restore a reference checkpoint into a VariableStore (and, when a Trainer with built optimiser state is given, its
    Adam moments and global step).  Returns {"loaded": [...], "missing": [...], "unused": [...], "global_step": int|None}."""
    import torch
    ck = read_tf_checkpoint(prefix, verify=verify)
    loaded, missing = [], []
    sd = {}
    for name, var in store.vars.items():
        src = ck.get(TOWER + name, ck.get(name))
        if src is None:
            missing.append(name)
            continue
        if tuple(src.shape) != tuple(var.shape):
            raise ValueError(f"{name}: checkpoint shape {tuple(src.shape)} != variable shape {tuple(var.shape)}")
        sd[name] = torch.from_numpy(np.array(src, dtype=np.float32))
        loaded.append(name)
    if strict and missing:
        raise KeyError(f"variables missing from the checkpoint: {missing[:8]}{' ...' if len(missing) > 8 else ''}")
    store.load_state_dict(sd)
    step = int(ck["global_step"]) if "global_step" in ck else None
    if trainer is not None:
        if step is not None:
            trainer.global_step = step
        slots = {k: v for k, v in ck.items() if k.endswith(("/Adam", "/Adam_1"))}

        def apply(flat):
            for name in flat.order:
                for slot, buf in (("Adam", flat.m), ("Adam_1", flat.v)):
                    src = slots.get(f"{TOWER}{name}/{slot}", slots.get(f"{name}/{slot}"))
                    if src is not None:
                        o, n = flat.offsets[name], store.vars[name].numel()
                        buf[o:o + n].copy_(torch.from_numpy(np.array(src, dtype=np.float32)).reshape(-1))

        if trainer.flat is not None:
            apply(trainer.flat)
        elif slots:
            trainer.on_flat_created = apply      # the flat optimiser state is laid out during the first step
    used = {TOWER + n for n in loaded} | set(loaded)
    unused = sorted(k for k in ck if k not in used and not k.endswith(("/Adam", "/Adam_1"))
                    and k not in ("global_step", "beta1_power", "beta2_power"))
    return {"loaded": loaded, "missing": missing, "unused": unused, "global_step": step}


def save_from_store(store, prefix: str, trainer=None, tower_scope: bool = True):
    """Write the variables (plus Adam slots / beta powers when a Trainer is given) under the names the reference's
    Saver uses.  `global_step` is ALWAYS written, as int32: the reference creates it with
    `tf.Variable(0, trainable=False, name="global_step")` (train.py:228, eval.py:131), eval.py restores
    `tf.train.Saver(tf.global_variables())` and train.py resumes through the Supervisor -- both need the key and the
    dtype to match.  inference.py restores through `import_meta_graph` and additionally needs a `.meta` graph file,
    which only TensorFlow can write."""
    pre = TOWER if tower_scope else ""
    if trainer is not None:
        trainer.sync_parameters()
    out = {pre + k: v.detach().cpu().numpy() for k, v in store.vars.items()}
    out["global_step"] = np.array(trainer.global_step if trainer is not None else 0, dtype=np.int32)
    if trainer is not None:
        t = max(trainer.global_step, 0)
        out["beta1_power"] = np.array(0.9 ** (t + 1), dtype=np.float32)     # AdamOptimizer's non-slot variables
        out["beta2_power"] = np.array(0.999 ** (t + 1), dtype=np.float32)
        flat = trainer.flat
        if flat is not None:
            tr = store.trainable()
            for name in flat.order:
                o, n = flat.offsets[name], tr[name].numel()
                out[f"{pre}{name}/Adam"] = flat.m[o:o + n].view(tr[name].shape).cpu().numpy()
                out[f"{pre}{name}/Adam_1"] = flat.v[o:o + n].view(tr[name].shape).cpu().numpy()
    write_tf_checkpoint(prefix, out)
    update_checkpoint_state(os.path.dirname(os.path.abspath(prefix)), prefix)


# ------------------------------------------------------------------------------------------------
# command line: python -m learnablepoolingmethods_b200.checkpoint {list|to-torch|from-torch} ...
# ------------------------------------------------------------------------------------------------
def _main(argv=None) -> int:
    import argparse
    ap = argparse.ArgumentParser(prog="python -m learnablepoolingmethods_b200.checkpoint",
                                 description="Inspect / convert TF-1.x V2 checkpoints of the reference trainer (no TensorFlow).")
    sub = ap.add_subparsers(dest="cmd", required=True)
    a = sub.add_parser("list", help="variables of a checkpoint (prefix, or a train_dir with a `checkpoint` state file)")
    a.add_argument("path")
    b = sub.add_parser("to-torch", help="checkpoint -> torch state dict (.pt) keyed by the VariableStore names")
    b.add_argument("path"); b.add_argument("out")
    b.add_argument("--keep-slots", action="store_true", help="also keep the Adam slots and global_step")
    c = sub.add_parser("from-torch", help="torch state dict (.pt) -> checkpoint the reference's eval.py / inference.py restore")
    c.add_argument("state"); c.add_argument("prefix")
    c.add_argument("--no-tower-scope", action="store_true", help="write bare names instead of tower/<name>")
    args = ap.parse_args(argv)

    def resolve(p):
        if os.path.isdir(p):
            q = latest_checkpoint(p)
            if q is None:
                raise SystemExit(f"{p}: no `checkpoint` state file")
            return q
        return p

    if args.cmd == "list":
        prefix = resolve(args.path)
        info = list_tf_checkpoint(prefix)
        total = 0
        for name, e in info.items():
            n = int(np.prod(e["shape"])) if e["shape"] else 1
            total += n
            print(f"{name:72s} {str(_DT.get(e['dtype'], e['dtype']).__name__ if e['dtype'] in _DT else e['dtype']):8s} {e['shape']}")
        print(f"{len(info)} tensors, {total} elements, {prefix}")
        return 0
    import torch
    if args.cmd == "to-torch":
        ck = read_tf_checkpoint(resolve(args.path))
        out = {}
        for k, v in ck.items():
            slot = k.endswith(("/Adam", "/Adam_1")) or k in ("global_step", "beta1_power", "beta2_power")
            if slot and not args.keep_slots:
                continue
            out[k[len(TOWER):] if k.startswith(TOWER) else k] = torch.from_numpy(np.array(v))
        torch.save(out, args.out)
        print(f"{len(out)} tensors -> {args.out}")
        return 0
    sd = torch.load(args.state, map_location="cpu")
    pre = "" if args.no_tower_scope else TOWER
    glob = ("global_step", "beta1_power", "beta2_power")
    tensors = {(k if k in glob else pre + k): v.detach().cpu().numpy() for k, v in sd.items()}
    # the reference's global_step variable is int32 (train.py:228); written even when the state dict has none
    tensors["global_step"] = np.array(int(tensors.get("global_step", 0)), dtype=np.int32)
    write_tf_checkpoint(args.prefix, tensors)
    update_checkpoint_state(os.path.dirname(os.path.abspath(args.prefix)), args.prefix)
    print(f"{len(sd)} tensors -> {args.prefix}.index / .data-00000-of-00001")
    return 0


if __name__ == "__main__":
    raise SystemExit(_main())
