"""sourmash signature (``.sig``) JSON files: the on-disk cache format of the sourmash method.

The reference stores one ``{md5}.sig`` per genome under ``cache/sourmash_k={k}_scaled={s}/``
(``pyani_plus/methods/sourmash.py:57-66``), written by ``sourmash scripts singlesketch``.  The format
is pinned key-for-key by the reference's ``tests/snakemake/test_sourmash_workflow.py:43-67`` against
the fixture files, so files written here are interchangeable with sourmash's.
"""

from __future__ import annotations

import hashlib
import json
from pathlib import Path

import numpy as np

SEED = 42


def sketch_md5sum(hashes: np.ndarray, ksize: int) -> str:
    """sourmash's sketch checksum: md5 of ``str(ksize)`` followed by every hash in decimal."""
    return hashlib.md5((str(ksize) + "".join(map(str, hashes.tolist()))).encode()).hexdigest()  # noqa: S324


def write_sig(path: Path, *, filename: str, name: str, ksize: int, max_hash: int, hashes: np.ndarray) -> None:
    """Write one single-sketch DNA signature file (same keys and key order as sourmash 4.8 / branchwater)."""
    mins = list(map(str, hashes.astype(np.uint64).tolist()))
    md5sum = hashlib.md5((str(ksize) + "".join(mins)).encode()).hexdigest()  # noqa: S324
    head = {
        "class": "sourmash_signature",
        "email": "",
        "hash_function": "0.murmur64",
        "filename": filename,
        "name": name,
        "license": "CC0",
    }
    sketch_head = {"num": 0, "ksize": ksize, "seed": SEED, "max_hash": max_hash}
    text = (
        "[{" + json.dumps(head, separators=(",", ":"))[1:-1]
        + ',"signatures":[{' + json.dumps(sketch_head, separators=(",", ":"))[1:-1]
        + ',"mins":[' + ",".join(mins) + "]"
        + ',"md5sum":"' + md5sum + '","molecule":"DNA"}],"version":0.4}]'
    )
    tmp = path.with_suffix(path.suffix + ".tmp")
    tmp.write_text(text)
    tmp.replace(path)  # never leave a half-written signature in the cache


def read_sig(path: Path, *, ksize: int | None = None) -> dict:
    """Read a signature file; returns name, ksize, max_hash, seed and the hashes as uint64.

    Picks the DNA sketch with the requested ``ksize`` when a file holds several.
    """
    data = json.loads(Path(path).read_text())
    if not isinstance(data, list) or not data:
        msg = f"{path} is not a sourmash signature file"
        raise ValueError(msg)
    outer = data[0]
    for sketch in outer.get("signatures", []):
        if sketch.get("molecule", "DNA") != "DNA":
            continue
        if ksize is not None and sketch.get("ksize") != ksize:
            continue
        hashes = np.asarray(sketch["mins"], dtype=np.uint64)
        return {
            "name": outer.get("name", ""),
            "filename": outer.get("filename", ""),
            "ksize": int(sketch["ksize"]),
            "seed": int(sketch.get("seed", SEED)),
            "max_hash": int(sketch["max_hash"]),
            "md5sum": sketch.get("md5sum", ""),
            "hashes": hashes,
        }
    msg = f"{path} holds no DNA sketch" + (f" with ksize={ksize}" if ksize is not None else "")
    raise ValueError(msg)
