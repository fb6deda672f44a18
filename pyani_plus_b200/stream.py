"""Host-side layout of the tiled base stream the sketch kernel consumes (see include/panib200.h).

Genome ``g`` owns tiles ``[tile_off[g], tile_off[g+1])`` of ``TILE`` base positions; its FASTA
records are written back to back with one invalid separator byte between records and the rest of
the last tile is invalid padding (always at least one byte), so that no k-mer window can span two
records (sourmash hashes each record on its own) or two genomes.  One extra all-invalid tile ends
the stream.  Invalid = any byte outside ``ACGTacgt``; we write ``N``.
"""

from __future__ import annotations

import numpy as np

TILE = 4096  # == PANIB_TILE_BASES
PAD = ord("N")


def genome_stream_length(records: list[bytes]) -> int:
    """Bases plus one separator between consecutive records."""
    if not records:
        return 0
    return sum(len(r) for r in records) + len(records) - 1


def plan_tiles(stream_lengths: list[int]) -> np.ndarray:
    """int64 tile offsets, one more entry than genomes."""
    tiles = np.asarray([length // TILE + 1 for length in stream_lengths], dtype=np.int64)
    off = np.zeros(len(stream_lengths) + 1, dtype=np.int64)
    np.cumsum(tiles, out=off[1:])
    return off


def stream_bytes(tile_off: np.ndarray) -> int:
    """Bytes of the ASCII stream including the trailing all-invalid tile."""
    return int(tile_off[-1] + 1) * TILE


def fill_ascii_stream(buf: np.ndarray, tile_off: np.ndarray, genomes: list[list[bytes]]) -> None:
    """Write the genomes' records into ``buf`` (uint8, ``stream_bytes(tile_off)`` long)."""
    if buf.dtype != np.uint8 or buf.ndim != 1 or buf.size != stream_bytes(tile_off):
        msg = "ASCII stream buffer has the wrong shape"
        raise ValueError(msg)
    buf[:] = PAD
    for g, records in enumerate(genomes):
        pos = int(tile_off[g]) * TILE
        for r, rec in enumerate(records):
            if r:
                pos += 1  # separator stays invalid
            n = len(rec)
            if n:
                buf[pos: pos + n] = rec if isinstance(rec, np.ndarray) else np.frombuffer(rec, dtype=np.uint8)
            pos += n
