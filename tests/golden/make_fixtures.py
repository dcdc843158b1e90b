#!/usr/bin/env python
"""Generate the committed golden fixtures from the reference tree (run in the build container only).

Reads /root/reference/cmd/testdata/{undeleteable-object.tgz,xl-meta-merge.zip,xl-meta-inline-notinline.zip}
— real bytes written by MinIO — verifies them IN FULL against the CPU oracle, and writes reduced
fixtures next to this script so the GPU box (which has no /root/reference) can replay them:

  rs75_fixture.npz      RS(7,5), 1 MiB blocks, object bucket/2 (5 MiB): for erasure blocks 0 and 4, the
                        first/last 4 KiB of every shard (pins parity rows + 3 zero pad bytes), the 60
                        frame digests, and two complete frames (one data, one parity; HH tail 5).
  inline_frames.json    [digest, shard] pairs cut from inline data of fixture xl.meta files
                        (HighwayHash goldens with many different tail lengths).
  inline_notinline.npz  the part.1 + xl.meta pair of cmd/erasure-object_test.go:1131-1184 (MD5 golden).

Nothing here is reference *source*; these are data fixtures (test vectors).
"""
import hashlib, io, json, os, sys, tarfile, zipfile
import msgpack
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle_lib as o  # noqa: E402

TD = "/root/reference/cmd/testdata"


def parse_xlmeta(buf):
    """-> (list of version dicts, inline map) ; format per cmd/xl-storage-format-v2.go:44-70,1178-1202"""
    if buf[:4] != b"XL2 ":
        return [], {}
    major, minor = int.from_bytes(buf[4:6], "little"), int.from_bytes(buf[6:8], "little")
    rest = buf[8:]
    versions, inline = [], {}
    if major == 1 and minor >= 1:
        up = msgpack.Unpacker(io.BytesIO(rest), raw=True, strict_map_key=False)
        blob = up.unpack()
        consumed = up.tell()
        if minor >= 2:
            up.unpack()  # crc
            consumed = up.tell()
        tail = rest[consumed:]
        if minor >= 3:
            bu = msgpack.Unpacker(io.BytesIO(blob), raw=True, strict_map_key=False)
            try:
                bu.unpack(); bu.unpack(); nv = bu.unpack()
                for _ in range(nv):
                    bu.unpack()
                    body = bu.unpack()
                    versions.append(msgpack.unpackb(body, raw=True, strict_map_key=False))
            except Exception:
                pass
        else:
            try:
                d = msgpack.unpackb(blob, raw=True, strict_map_key=False)
                versions = d.get(b"Versions", [])
            except Exception:
                pass
        if len(tail) > 1 and tail[0] == 1:
            try:
                inline = msgpack.unpackb(tail[1:], raw=True, strict_map_key=False)
            except Exception:
                inline = {}
    return versions, inline


def collect_inline(name, buf, out):
    vers, inline = parse_xlmeta(buf)
    for key, val in (inline or {}).items():
        if not isinstance(val, (bytes, bytearray)) or len(val) <= 32:
            continue
        dig, shard = bytes(val[:32]), bytes(val[32:])
        # single-frame inline objects only (shard file < one erasure block)
        if o.hh256(shard) == dig:
            out.append((name, dig.hex(), shard))


def main():
    frames = []
    # ---------------- undeleteable-object.tgz : RS(7,5) ----------------
    tf = tarfile.open(os.path.join(TD, "undeleteable-object.tgz"))
    parts, metas = {}, {}
    for mem in tf.getmembers():
        if not mem.isfile():
            continue
        data = tf.extractfile(mem).read()
        segs = mem.name.split("/")
        drive = next((s for s in segs if s.startswith("xl") and s[2:].isdigit()), None)
        if mem.name.endswith("part.1") and drive:
            parts[int(drive[2:])] = data
        if mem.name.endswith("xl.meta"):
            collect_inline(mem.name, data, frames)
            if "/bucket/2/xl.meta" in mem.name and drive:
                metas[int(drive[2:])] = data
    k, m, bs, size = 7, 5, 1 << 20, 5 * (1 << 20)
    dist = None
    for d, buf in metas.items():
        for v in parse_xlmeta(buf)[0]:
            ob = v.get(b"V2Obj")
            if ob and ob.get(b"EcM") == 7 and ob.get(b"Size") == size:
                dist = list(ob[b"EcDist"]); assert ob[b"EcN"] == 5 and ob[b"EcBSize"] == bs
    assert dist == [5, 6, 7, 8, 9, 10, 11, 12, 1, 2, 3, 4], dist
    S = o.shard_size(bs, k)
    assert S == 149797
    files = [None] * 12
    for d in range(1, 13):
        files[dist[d - 1] - 1] = np.frombuffer(parts[d], dtype=np.uint8)
        assert len(parts[d]) == 5 * (32 + S)
    # full verification against the oracle: every frame digest, every parity byte, padding
    digests = np.zeros((5, 12, 32), dtype=np.uint8)
    obj = bytearray()
    for b in range(5):
        shards = []
        for i in range(12):
            fr = files[i][b * (32 + S):(b + 1) * (32 + S)]
            assert o.hh256(fr[32:]) == fr[:32].tobytes(), (b, i)
            digests[b, i] = fr[:32]
            shards.append(fr[32:])
        blockdata = np.concatenate(shards[:k])[:bs]
        assert not np.concatenate(shards[:k])[bs:].any()  # 3 zero pad bytes
        enc = o.encode_data(k, m, blockdata)
        for i in range(12):
            assert np.array_equal(enc[i], shards[i]), (b, i)
        obj += blockdata.tobytes()
    # the oracle's own whole-object driver must reproduce the 12 part.1 files byte for byte
    ofiles, _ = o.erasure_encode(k, m, bs, o.HIGHWAYHASH256S, bytes(obj))
    for i in range(12):
        assert np.array_equal(ofiles[i], files[i])
    W = 4096
    head = np.stack([[files[i][b * (32 + S) + 32: b * (32 + S) + 32 + W] for i in range(12)] for b in (0, 4)])
    tail = np.stack([[files[i][(b + 1) * (32 + S) - W:(b + 1) * (32 + S)] for i in range(12)] for b in (0, 4)])
    np.savez_compressed(os.path.join(HERE, "rs75_fixture.npz"), k=k, m=m, block_size=bs, shard_size=S,
                        head=head, tail=tail, digests=digests,
                        frame_data=files[1][:32 + S].copy(), frame_parity=files[9][2 * (32 + S):3 * (32 + S)].copy(),
                        object_md5=np.frombuffer(hashlib.md5(obj).digest(), dtype=np.uint8))
    print("rs75: 60 frames + 25 parity blocks verified against oracle; object md5", hashlib.md5(obj).hexdigest())

    # ---------------- zips ----------------
    for zn in ("xl-meta-merge.zip", "xl-meta-consist.zip", "xl-meta-inline-notinline.zip"):
        zf = zipfile.ZipFile(os.path.join(TD, zn))
        for nm in zf.namelist():
            if nm.endswith("xl.meta"):
                collect_inline(zn + ":" + nm, zf.read(nm), frames)
    for nm in ("xl.meta", "xl-many-parts.meta"):
        collect_inline(nm, open(os.path.join(TD, nm), "rb").read(), frames)

    # keep a bounded, diverse set: up to 3 frames per (len mod 32), shards <= 20000 bytes
    by_tail, seen = {}, set()
    for name, dig, shard in sorted(frames, key=lambda t: len(t[2])):
        if len(shard) > 20000 or dig in seen:
            continue
        seen.add(dig)
        by_tail.setdefault(len(shard) % 32, [])
        if len(by_tail[len(shard) % 32]) < 3:
            by_tail[len(shard) % 32].append({"src": name, "len": len(shard), "digest": dig, "shard": shard.hex()})
    out = [f for t in sorted(by_tail) for f in by_tail[t]]
    json.dump(out, open(os.path.join(HERE, "inline_frames.json"), "w"), indent=0)
    print("inline frames:", len(out), "tails pinned:", sorted(by_tail))

    # ---------------- xl-meta-inline-notinline.zip (erasure-object_test.go:1131-1184) ----------------
    zf = zipfile.ZipFile(os.path.join(TD, "xl-meta-inline-notinline.zip"))
    names = zf.namelist()
    part = next(n for n in names if n.endswith("part.1"))
    meta2 = [n for n in names if n.endswith("xl.meta")]
    np.savez_compressed(os.path.join(HERE, "inline_notinline.npz"),
                        part1=np.frombuffer(zf.read(part), dtype=np.uint8),
                        **{"meta_" + n.split("/")[0]: np.frombuffer(zf.read(n), dtype=np.uint8) for n in meta2})
    print("inline-notinline:", part, meta2)


if __name__ == "__main__":
    main()
