"""sourmash signature (``.sig``) JSON files: the on-disk cache format of the sourmash method.

The reference stores one ``{md5}.sig`` per genome under ``cache/sourmash_k={k}_scaled={s}/``
(``pyani_plus/methods/sourmash.py:57-66``), written by ``sourmash scripts singlesketch``.  The format
is pinned key-for-key by the reference's ``tests/snakemake/test_sourmash_workflow.py:43-67`` against
the fixture files, so files written here are interchangeable with sourmash's.
"""

from __future__ import annotations

import hashlib
import json
import os
import struct
from pathlib import Path

import numpy as np

SEED = 42

# Binary side-cache: ``{md5}.sig.u64`` next to ``{md5}.sig``.  The JSON stays the interchange format (and the
# only thing other tools read); the side file holds the same hashes as raw little-endian uint64 behind a
# 48-byte header that ties it to ONE version of the JSON file (its size and mtime), so an externally replaced
# or edited ``.sig`` silently invalidates it.  Reading 10,000 sketches: ~0.4 GB of binary instead of
# ~0.9 GB of decimal text through ``json.loads``.
SIDE_SUFFIX = ".u64"
_SIDE_MAGIC = b"PANIBSK1"
_SIDE_HEADER = struct.Struct("<8sIIQQQQ")  # magic, ksize, seed, max_hash, count, sig_size, sig_mtime_ns


def _mins_text(hashes: np.ndarray) -> bytes:
    """The hashes in decimal, comma-separated (C formatter of the library: ``panib_format_u64``)."""
    from pyani_plus_b200 import engine  # noqa: PLC0415

    return engine.format_u64(hashes, b",")


def sketch_md5sum(hashes: np.ndarray, ksize: int) -> str:
    """sourmash's sketch checksum: md5 of ``str(ksize)`` followed by every hash in decimal."""
    return hashlib.md5(str(ksize).encode() + _mins_text(hashes).replace(b",", b"")).hexdigest()  # noqa: S324


def write_sig(path: Path, *, filename: str, name: str, ksize: int, max_hash: int, hashes: np.ndarray) -> None:
    """Write one single-sketch DNA signature file (same keys and key order as sourmash 4.8 / branchwater)."""
    mins = _mins_text(hashes)
    md5sum = hashlib.md5(str(ksize).encode() + mins.replace(b",", b"")).hexdigest()  # noqa: S324
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
        b"[{" + json.dumps(head, separators=(",", ":"))[1:-1].encode()
        + b',"signatures":[{' + json.dumps(sketch_head, separators=(",", ":"))[1:-1].encode()
        + b',"mins":[' + mins + b"]"
        + b',"md5sum":"' + md5sum.encode() + b'","molecule":"DNA"}],"version":0.4}]'
    )
    tmp = path.with_suffix(path.suffix + ".tmp")
    tmp.write_bytes(text)
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


def side_cache_path(sig_path: Path) -> Path:
    return sig_path.with_name(sig_path.name + SIDE_SUFFIX)


def write_side_cache(sig_path: Path, sig: dict) -> None:
    """Write the binary twin of ``sig_path`` (best effort: a read-only cache directory is not an error)."""
    try:
        st = sig_path.stat()
        hashes = np.ascontiguousarray(sig["hashes"], dtype="<u8")
        side = side_cache_path(sig_path)
        tmp = side.with_name(side.name + f".tmp{os.getpid()}")
        with tmp.open("wb") as handle:
            handle.write(_SIDE_HEADER.pack(_SIDE_MAGIC, int(sig["ksize"]), int(sig.get("seed", SEED)),
                                           int(sig["max_hash"]), int(hashes.size), st.st_size, st.st_mtime_ns))
            handle.write(hashes.tobytes())
        tmp.replace(side)
    except OSError:
        pass


def read_side_cache(sig_path: Path, st: os.stat_result | None = None) -> dict | None:
    """The sketch from the binary side-cache, or None when it is absent, truncated, or belongs to another
    version of the ``.sig`` file (size / mtime differ)."""
    side = side_cache_path(sig_path)
    try:
        raw = side.read_bytes()
    except OSError:
        return None
    if len(raw) < _SIDE_HEADER.size:
        return None
    magic, ksize, seed, max_hash, count, sig_size, sig_mtime = _SIDE_HEADER.unpack_from(raw)
    st = st or sig_path.stat()
    if magic != _SIDE_MAGIC or sig_size != st.st_size or sig_mtime != st.st_mtime_ns or \
            len(raw) != _SIDE_HEADER.size + 8 * count:
        return None
    hashes = np.frombuffer(raw, dtype="<u8", offset=_SIDE_HEADER.size).astype(np.uint64, copy=False)
    return {"name": sig_path.stem, "filename": "", "ksize": int(ksize), "seed": int(seed), "max_hash": int(max_hash),
            "md5sum": "", "hashes": hashes}
