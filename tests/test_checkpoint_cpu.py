"""TF-1.x V2 checkpoint (tensor bundle) reader / writer, model_flags.json, name mapping.  No TensorFlow
here: the format is checked through its own invariants (CRC-32C known answers, masked CRCs, table magic, prefix-compressed
blocks with restart points), a hand-assembled byte-level fixture, and writer -> reader round trips."""
import json
import os
import struct

import numpy as np
import pytest

from learnablepoolingmethods_b200 import checkpoint as ck


def test_crc32c_known_answers():
    # RFC 3720 B.4 test vectors + the classic check value
    assert ck.crc32c(b"123456789") == 0xE3069283
    assert ck.crc32c(bytes(32)) == 0x8A9136AA
    assert ck.crc32c(bytes([0xFF] * 32)) == 0x62A8AB43
    assert ck.crc32c(bytes(range(32))) == 0x46DD794E
    assert ck.crc32c(b"56789", ck.crc32c(b"1234")) == 0xE3069283            # continuation
    assert ck._py_crc32c(0, b"123456789") == 0xE3069283                     # pure-Python fallback agrees
    a = np.arange(1000, dtype=np.float32)
    assert ck.crc32c(a) == ck._py_crc32c(0, a.tobytes())
    # crc32c::Mask: rotate right 15, add 0xa282ead8 (leveldb / tensorflow crc32c.h)
    assert ck.mask_crc(0) == 0xA282EAD8
    assert ck.mask_crc(0x00008000) == (1 + 0xA282EAD8) & 0xFFFFFFFF


def test_reader_on_hand_assembled_bundle(tmp_path):
    """A two-tensor bundle built byte by byte from the published layout (not with the writer under test)."""
    w = np.array([[1.5, -2.0, 3.25], [0.0, 4.0, -8.0]], dtype="<f4")
    step = np.array(1234, dtype="<i8")
    data = step.tobytes() + w.tobytes()
    prefix = str(tmp_path / "model.ckpt-1234")
    open(prefix + ".data-00000-of-00001", "wb").write(data)

    def vi(n):
        return ck._put_varint(n)

    def entry(dtype, dims, offset, size, arr_bytes):
        shape = b"".join(b"\x12" + vi(len(d)) + d for d in (b"\x08" + vi(s) for s in dims))       # dim { size }
        m = b"\x08" + vi(dtype) + b"\x12" + vi(len(shape)) + shape
        if offset:
            m += b"\x20" + vi(offset)
        return m + b"\x28" + vi(size) + b"\x35" + struct.pack("<I", ck.mask_crc(ck._py_crc32c(0, arr_bytes)))

    kv = [(b"", b"\x08\x01\x1a\x02\x08\x01"),                                         # header: 1 shard, producer 1
          (b"global_step", entry(9, [], 0, 8, step.tobytes())),
          (b"tower/w", entry(1, [2, 3], 8, 24, w.tobytes()))]
    # one data block, no key sharing, a single restart point
    blk = b"".join(vi(0) + vi(len(k)) + vi(len(v)) + k + v for k, v in kv) + struct.pack("<II", 0, 1)

    def with_trailer(b):
        return b + b"\x00" + struct.pack("<I", ck.mask_crc(ck._py_crc32c(0, b + b"\x00")))

    out = with_trailer(blk)
    meta_off = len(out)
    meta = struct.pack("<II", 0, 1)
    out += with_trailer(meta)
    handle = vi(0) + vi(len(blk))
    idx = vi(0) + vi(len(b"tower/w")) + vi(len(handle)) + b"tower/w" + handle + struct.pack("<II", 0, 1)
    idx_off = len(out)
    out += with_trailer(idx)
    footer = vi(meta_off) + vi(len(meta)) + vi(idx_off) + vi(len(idx))
    out += footer + b"\x00" * (40 - len(footer)) + struct.pack("<Q", 0xDB4775248B80FB57)
    open(prefix + ".index", "wb").write(out)

    got = ck.read_tf_checkpoint(prefix)
    assert set(got) == {"global_step", "tower/w"}
    assert got["global_step"].dtype == np.int64 and int(got["global_step"]) == 1234
    np.testing.assert_array_equal(got["tower/w"], w)
    info = ck.list_tf_checkpoint(prefix)
    assert info["tower/w"]["shape"] == [2, 3] and info["tower/w"]["offset"] == 8 and info["tower/w"]["size"] == 24
    # corruption is detected: a flipped bit in the index, and (native CRC) in the tensor data
    bad = bytearray(out); bad[10] ^= 1
    open(prefix + ".index", "wb").write(bytes(bad))
    with pytest.raises(ValueError):
        ck.read_tf_checkpoint(prefix)
    open(prefix + ".index", "wb").write(out)
    if ck.have_native_crc():
        d2 = bytearray(data); d2[12] ^= 0x40
        open(prefix + ".data-00000-of-00001", "wb").write(bytes(d2))
        with pytest.raises(ValueError, match="checksum"):
            ck.read_tf_checkpoint(prefix)
    open(prefix + ".index", "wb").write(b"not a table" * 8)
    with pytest.raises(ValueError, match="magic"):
        ck.read_tf_checkpoint(prefix)


def test_round_trip_many_variables(tmp_path):
    """Enough keys for several 4 KB index blocks, shared prefixes, restart points, scalars, all dtypes used."""
    rng = np.random.RandomState(0)
    tensors = {"global_step": np.array(77, dtype=np.int64), "beta1_power": np.array(0.9 ** 78, dtype=np.float32)}
    for i in range(300):
        shape = [(), (3,), (4, 5), (1, 6, 2)][i % 4]
        tensors[f"tower/video_VLAD/cluster_attention/layer_{i:03d}/kernel"] = np.asarray(rng.randn(*shape), dtype=np.float32)
        tensors[f"tower/video_VLAD/cluster_attention/layer_{i:03d}/kernel/Adam"] = np.asarray(rng.randn(*shape), dtype=np.float32)
    tensors["flags/bool"] = np.array([True, False, True])
    tensors["half"] = rng.randn(7).astype(np.float16)
    tensors["i32"] = rng.randint(-5, 5, size=(2, 2)).astype(np.int32)
    prefix = str(tmp_path / "sub" / "model.ckpt-77")
    ck.write_tf_checkpoint(prefix, tensors)
    raw = open(prefix + ".index", "rb").read()
    assert struct.unpack("<Q", raw[-8:])[0] == 0xDB4775248B80FB57 and len(raw) > 3 * 4096
    got = ck.read_tf_checkpoint(prefix)
    assert set(got) == set(tensors)
    for k, v in tensors.items():
        assert got[k].dtype == v.dtype and got[k].shape == v.shape, k
        np.testing.assert_array_equal(got[k], v)
    # keys come back in bytewise order (the table is sorted), offsets tile the data file exactly
    info = ck.list_tf_checkpoint(prefix)
    assert list(info) == sorted(info, key=lambda s: s.encode())
    assert sum(e["size"] for e in info.values()) == os.path.getsize(prefix + ".data-00000-of-00001")
    sub = ck.read_tf_checkpoint(prefix, names=["half", "i32"])
    assert set(sub) == {"half", "i32"}
    with pytest.raises(KeyError):
        ck.read_tf_checkpoint(prefix, names=["nope"])


def test_model_flags_and_checkpoint_state(tmp_path):
    d = str(tmp_path)
    f = ck.model_flags(model="NetVladV1")
    ck.write_model_flags(d, f)
    assert json.load(open(os.path.join(d, "model_flags.json"))) == {
        "model": "NetVladV1", "feature_sizes": "1024,128", "feature_names": "rgb,audio", "frame_features": True,
        "label_loss": "CrossEntropyLoss"}                                    # train.py:390-396
    ck.write_model_flags(d, f)                                               # same flags: accepted
    with pytest.raises(ValueError):
        ck.write_model_flags(d, ck.model_flags(model="NetVladV2"))           # train.py:398-407 exits
    assert ck.read_model_flags(d)["model"] == "NetVladV1"
    assert ck.latest_checkpoint(d) is None
    ck.update_checkpoint_state(d, os.path.join(d, "model.ckpt-10"))
    assert ck.latest_checkpoint(d) == os.path.join(d, "model.ckpt-10")


def test_store_round_trip_with_tower_scope(tmp_path):
    """VariableStore -> checkpoint (tower/ names, as the reference's Saver writes them) -> fresh store."""
    import torch
    from learnablepoolingmethods_b200 import variables
    from learnablepoolingmethods_b200.engine import NetVladConfig, NetVladEngine
    cfg = NetVladConfig(model="NetVladV1", iterations=16, cluster_size=8, hidden_size=32, vocab_size=50, rgb_dim=64, audio_dim=16,
                        rgb_heads=4, audio_heads=2)
    a = variables.VariableStore("cpu", seed=1)
    NetVladEngine.build_variables(type("E", (), {"cfg": cfg, "store": a, "_dense_vars": NetVladEngine._dense_vars,
                                                  "_ln_vars": NetVladEngine._ln_vars, "wc_suffix": NetVladEngine.wc_suffix})())
    prefix = str(tmp_path / "model.ckpt-5")
    ck.save_from_store(a, prefix)
    names = ck.list_tf_checkpoint(prefix)
    assert "tower/video_VLAD/cluster_weights2" in names and names["tower/video_VLAD/cluster_weights2"]["shape"] == [1, 64, 8]
    assert "tower/audio_attention/filter_outputencode2/kernel" in names and "tower/gates/weights" in names
    # the reference's Saver(tf.global_variables()) needs global_step: always present, int32 (train.py:228, eval.py:131)
    assert names["global_step"]["shape"] == [] and ck.read_tf_checkpoint(prefix)["global_step"].dtype == np.int32
    b = variables.VariableStore("cpu", seed=2)
    for k, v in a.vars.items():
        b.vars[k] = torch.zeros_like(v)
    rep = ck.load_into_store(b, prefix)
    assert not rep["missing"] and not rep["unused"] and len(rep["loaded"]) == len(a.vars)
    for k in a.vars:
        assert torch.equal(a.vars[k], b.vars[k]), k
    b.vars["extra/var"] = torch.zeros(3)
    with pytest.raises(KeyError):
        ck.load_into_store(b, prefix)
    assert ck.load_into_store(b, prefix, strict=False)["missing"] == ["extra/var"]
    b.vars["hidden1_biases"] = torch.zeros(5)
    with pytest.raises(ValueError, match="shape"):
        ck.load_into_store(b, prefix, strict=False)


def test_command_line_round_trip(tmp_path, capsys):
    """python -m learnablepoolingmethods_b200.checkpoint from-torch / list / to-torch."""
    import torch
    sd = {"input_bn/gamma": torch.rand(8), "video_VLAD/cluster_weights": torch.randn(8, 4), "global_step": torch.tensor(12)}
    pt = str(tmp_path / "state.pt")
    torch.save(sd, pt)
    prefix = str(tmp_path / "train_dir" / "model.ckpt-12")
    assert ck._main(["from-torch", pt, prefix]) == 0
    names = ck.list_tf_checkpoint(prefix)
    assert set(names) == {"tower/input_bn/gamma", "tower/video_VLAD/cluster_weights", "global_step"}
    assert ck._main(["list", str(tmp_path / "train_dir")]) == 0          # a train_dir resolves through its `checkpoint` file
    assert "tower/video_VLAD/cluster_weights" in capsys.readouterr().out
    back = str(tmp_path / "back.pt")
    assert ck._main(["to-torch", prefix, back]) == 0
    got = torch.load(back)
    assert set(got) == {"input_bn/gamma", "video_VLAD/cluster_weights"}  # slots / global_step dropped by default
    assert torch.equal(got["video_VLAD/cluster_weights"], sd["video_VLAD/cluster_weights"])
    assert ck._main(["to-torch", prefix, back, "--keep-slots"]) == 0 and int(torch.load(back)["global_step"]) == 12
    assert ck.read_tf_checkpoint(prefix)["global_step"].dtype == np.int32
