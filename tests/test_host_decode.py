"""CPU: the C++ KVEvents msgpack decoder (host mirror, no GPU needed) vs the Python oracle's typed restatement of
processEvent / getHashAsUint64 (pkg/kvcache/kvevents/pool.go:177-244, :343-367; events.go:38-96)."""
import random
import struct

import msgpack
import numpy as np
import pytest

from kvidx.host import HostIndexer
from oracle import kvoracle as ko


def canon_oracle(payload):
    out = []
    for e in ko.decode_event_batch(payload):
        if isinstance(e, ko.BlockStored):
            out.append(("S", [h & ko.MASK64 for h in e.block_hashes], e.parent_block_hash, list(e.token_ids),
                        ko.medium_tier(e.medium)))
        elif isinstance(e, ko.BlockRemoved):
            out.append(("R", [h & ko.MASK64 for h in e.block_hashes], None, [], ko.medium_tier(e.medium)))
    return out


def canon_host(h, payload):
    ev, hs, tk = h.decode("pod-x", "model-y", payload)
    out = []
    for e in ev:
        hashes = [int(x) for x in hs[e["hash_off"]: e["hash_off"] + e["n_hashes"]]]
        tier = int(e["podtier"]) & 15
        assert (int(e["podtier"]) >> 4) == h.pod_id("pod-x")
        if e["op"] == 0:
            toks = [int(x) for x in tk[e["tok_off"]: e["tok_off"] + e["n_tokens"]]]
            out.append(("S", hashes, int(e["parent_hash"]) if e["has_parent"] else None, toks, tier))
        else:
            out.append(("R", hashes, None, [], tier))
    return out


def same(h, payload):
    a, b = canon_oracle(payload), canon_host(h, payload)
    a = [(x[0], x[1], x[2], x[3], h.tier_id(x[4])) for x in a]
    assert a == b, (payload.hex()[:200], a, b)
    return a


H64 = 2 ** 63 + 12345


@pytest.fixture(scope="module")
def host():
    return HostIndexer(no_device=True)


def test_well_formed_batches(host):
    p = msgpack.packb([1723.5, [["BlockStored", [H64, H64 + 1], None, list(range(32)), 16, None, "GPU"],
                                ["BlockStored", [H64 + 2], H64 + 1, [70000 + i for i in range(16)], 16],
                                ["BlockRemoved", [H64], "cpu"], ["AllBlocksCleared"],
                                ["BlockStored", [b"\x01" * 32, b"\xff\xee"], b"\x00" * 7 + b"\x09" * 3, [5] * 16, 16, 3, "Disk"]], 0])
    got = same(host, p)
    assert [g[0] for g in got] == ["S", "S", "R", "S"]
    assert got[0][4] == host.tier_id("gpu") and got[2][4] == host.tier_id("cpu") and got[3][4] == host.tier_id("disk")
    assert got[3][1] == [int.from_bytes(b"\x01" * 8, "big"), 0xFFEE] and got[3][2] == int.from_bytes((b"\x00" * 7 + b"\x09" * 3)[-8:], "big")


def test_hash_width_rules(host):
    """Only uint64- / int64-coded integers and byte slices are hashes (getHashAsUint64 on DecodeInterface's types)."""
    def ints(code, fmt, v):
        return bytes([code]) + struct.pack(fmt, v)
    arr = b"\x98\x08" if False else None
    items = [b"\x05", ints(0xcc, ">B", 200), ints(0xcd, ">H", 60000), ints(0xce, ">I", 4000000000), ints(0xcf, ">Q", H64),
             ints(0xd0, ">b", -5), ints(0xd1, ">h", -500), ints(0xd2, ">i", -70000), ints(0xd3, ">q", -7), b"\xa3abc", b"\xc4\x00", b"\xc4\x03\x01\x02\x03",
             b"\xc0", b"\xcb" + struct.pack(">d", 1.5), b"\x92\x01\x02"]
    hashes = bytes([0xdc]) + struct.pack(">H", len(items)) + b"".join(items)
    ev = b"\x95" + msgpack.packb("BlockStored") + hashes + b"\xc0" + msgpack.packb(list(range(16))) + b"\x10"
    payload = b"\x92" + msgpack.packb(1.0) + b"\x91" + ev
    got = same(host, payload)
    assert got[0][1] == [H64, (-7) & ko.MASK64, 0x010203]
    # parent of an unsupported width skips the whole event; nil parent is fine
    ev2 = b"\x95" + msgpack.packb("BlockStored") + msgpack.packb([H64]) + b"\x07" + msgpack.packb(list(range(16))) + b"\x10"
    assert same(host, b"\x92" + msgpack.packb(1.0) + b"\x91" + ev2) == []


def test_malformed_inputs(host):
    good = ["BlockStored", [H64], None, list(range(16)), 16]
    cases = [b"", b"\xc1", msgpack.packb({"a": 1}), msgpack.packb("str"), msgpack.packb([1.0]), msgpack.packb([1.0, None]),
             msgpack.packb([1.0, "notarray"]), msgpack.packb(["ts", [good]]), msgpack.packb([1, [good], "rank"]),
             msgpack.packb([1.0, [good, 5, [], [5], ["BlockStored"], ["BlockStored", "x"], ["BlockStored", [H64], None, "toks"],
                                  ["BlockStored", [H64], None, [1.5]], ["BlockRemoved"], ["BlockRemoved", [H64], 7], ["Other", 1]]]),
             msgpack.packb([1.0, [good]])[:-3], msgpack.packb([1.0, [good], 0, "extra", [1, 2]])]
    for c in cases:
        same(host, c)
    assert len(same(host, cases[-1])) == 1 and same(host, cases[-2]) == []


def test_fuzz_differential():
    rnd = random.Random(7)
    host = HostIndexer(no_device=True)
    base = msgpack.packb([1723.5, [["BlockStored", [H64, b"\x07" * 9], H64 - 1, list(range(100, 132)), 16, None, "GPU"],
                                   ["BlockRemoved", [H64, H64 + 5], "CPU"], ["BlockStored", [H64 + 9], None, [2 ** 31 + i for i in range(16)], 16]], 1])
    for it in range(3000):
        if it % 8 == 0:            # mutated Medium strings intern as new tiers; the id space is 4 bits (16 tier names)
            host = HostIndexer(no_device=True)
        b = bytearray(base)
        for _ in range(rnd.randrange(1, 4)):
            r = rnd.random()
            pos = rnd.randrange(len(b))
            if r < 0.6:
                b[pos] = rnd.randrange(256)
            elif r < 0.8:
                del b[pos]
            else:
                b.insert(pos, rnd.randrange(256))
        if any(isinstance(e, (ko.BlockStored, ko.BlockRemoved)) and e.medium is not None and "\x00" in e.medium
               for e in ko.decode_event_batch(bytes(b))):
            continue               # a NUL inside a Medium name cannot be passed to kvhost_tier_id (C string) for the comparison
        same(host, bytes(b))


def test_queue_sharding(host):
    for pod in ("pod-1", "10.0.0.7:8000", "x" * 40):
        assert host.queue_index(pod) == ko.fnv32a(pod.encode()) % 4
