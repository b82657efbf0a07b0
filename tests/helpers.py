"""Shared test plumbing: golden loading, string interning, scenario replay on id-level backends."""
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden")
M64 = (1 << 64) - 1
EVENT_DTYPE = np.dtype([("op", "u1"), ("has_parent", "u1"), ("podtier", "<u2"), ("model", "<u4"),
                        ("parent_hash", "<u8"), ("hash_off", "<u8"), ("tok_off", "<u8"),
                        ("n_hashes", "<u4"), ("n_tokens", "<u4")])


def golden(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


def csr(prompts):
    """list of token lists -> (tok uint32, off int64)."""
    off = np.zeros(len(prompts) + 1, np.int64)
    for i, p in enumerate(prompts):
        off[i + 1] = off[i] + len(p)
    tok = np.zeros(max(int(off[-1]), 1), np.uint32)
    for i, p in enumerate(prompts):
        tok[off[i]:off[i + 1]] = np.asarray(p, np.uint64).astype(np.uint32)
    return tok[:int(off[-1])] if off[-1] else tok[:0], off


class Interner:
    def __init__(self, names=()):
        self.ids = {}
        for n in names:
            self.id(n)

    def id(self, name):
        if name not in self.ids:
            self.ids[name] = len(self.ids)
        return self.ids[name]


def filter_mask(pod_ids, words):
    m = np.zeros(words, np.uint64)
    for p in pod_ids:
        m[p // 64] |= np.uint64(1) << np.uint64(p % 64)
    return m


def scenario_events(sc, pods: Interner, tiers: Interner, model_id=0):
    """scenario_small.json events -> (kvidx_event_t array, hashes, tokens) in arrival order."""
    ev, hashes, tokens = [], [], []
    for e in sc["events"]:
        if e["type"] == "AllBlocksCleared":
            continue                         # no-op in the reference (kvevents/pool.go:332-333); dropped on the host
        med = e.get("medium")
        tier = tiers.id(med.lower() if med is not None else "gpu")
        r = np.zeros((), EVENT_DTYPE)
        r["op"] = 0 if e["type"] == "BlockStored" else 1
        r["podtier"] = (pods.id(e["pod"]) << 4) | tier
        r["model"] = model_id
        r["hash_off"] = len(hashes)
        r["n_hashes"] = len(e["hashes"])
        hashes.extend(e["hashes"])
        if e["type"] == "BlockStored":
            r["has_parent"] = 0 if e["parent"] is None else 1
            r["parent_hash"] = e["parent"] or 0
            r["tok_off"] = len(tokens)
            r["n_tokens"] = len(e["tokens"])
            tokens.extend(e["tokens"])
        ev.append(r)
    return (np.array(ev, EVENT_DTYPE), np.array(hashes, np.uint64), np.array(tokens, np.uint64).astype(np.uint32))


def dense_from_map(scores, pods: Interner, P):
    row = np.full(P, -1.0)
    if scores:
        for p, s in scores.items():
            row[pods.ids[p]] = s
    return row
