"""ctypes binding of ``libpanib200.so`` (C ABI: ``include/panib200.h``) plus device-buffer plumbing.

PyTorch is used only to own device / pinned memory and the CUDA stream; every kernel is ours.
There is deliberately NO CPU fallback: if the shared library is missing or no CUDA device is
visible, constructing :class:`Engine` raises.

What the engine replaces in the reference: the two external commands of
``pyani_plus/methods/sourmash.py`` -- ``sourmash scripts singlesketch`` (:67-83) and
``sourmash scripts manysearch`` (:184-200).
"""

from __future__ import annotations

import ctypes
import os
from dataclasses import dataclass
from pathlib import Path

import numpy as np

from pyani_plus_b200 import stream as _stream

_PKG = Path(__file__).resolve().parent
LIB_PATH = _PKG / "libpanib200.so"

PROGRAM = "panib200"  # recorded as Configuration.program (the reference records "sourmash")

ST_BUCKET_OVERFLOW = 1
ST_SEGMENT_OVERFLOW = 2
ST_INDEX_OVERFLOW = 4

_vp = ctypes.c_void_p
_i64 = ctypes.c_int64
_u64 = ctypes.c_uint64
_i32 = ctypes.c_int
_u32 = ctypes.c_uint32


class EngineError(RuntimeError):
    """The CUDA engine is unavailable or a call into it failed."""


_lib: ctypes.CDLL | None = None
_WORKSPACES: dict = {}  # device index -> (survivor scratch tensor registered with the library, frozen by a graph)


def load_library() -> ctypes.CDLL:
    """Load libpanib200.so and declare the prototypes of include/panib200.h."""
    global _lib  # noqa: PLW0603
    if _lib is not None:
        return _lib
    import os  # noqa: PLC0415

    lib_path = Path(os.environ.get("PANIB200_LIB", LIB_PATH))  # override = kernel-variant experiments
    if not lib_path.is_file():
        msg = (
            f"{lib_path} is missing - build it with `python -c 'import __graft_entry__ as g; g.build()'`."
            " There is no CPU fallback for the sourmash path."
        )
        raise EngineError(msg)
    L = ctypes.CDLL(str(lib_path))
    L.panib_version.restype = ctypes.c_char_p
    L.panib_last_error.restype = _i32
    L.panib_last_error.argtypes = [ctypes.c_char_p, ctypes.c_size_t]
    L.panib_device_count.restype = _i32
    L.panib_launch_count.restype = _u64
    L.panib_max_hash.restype = _u64
    L.panib_max_hash.argtypes = [_u64]
    L.panib_plan_buckets.restype = _i32
    L.panib_plan_buckets.argtypes = [_i64, _u64, ctypes.c_double, ctypes.POINTER(ctypes.c_int32),
                                     ctypes.POINTER(_u64)]
    L.panib_pack_ascii.restype = _i32
    L.panib_pack_ascii.argtypes = [_vp, _i64, _vp, _vp, _vp]
    sk_args = [_vp, _vp, _vp, _i64, _i64, _i32, _u32, _u64, _vp, _vp, _vp, _i64]
    L.panib_sketch_stream.restype = _i32
    L.panib_sketch_stream.argtypes = [*sk_args, _vp, _vp, _vp, _vp]
    L.panib_sketch_hash_only.restype = _i32
    L.panib_sketch_hash_only.argtypes = [*sk_args, _vp, _vp, _vp]
    L.panib_sketch_finalize.restype = _i32
    L.panib_sketch_finalize.argtypes = [_vp, _i64, _i64, _vp, _vp, _vp, _vp, _vp]
    L.panib_sketch_ascii_host.restype = _i32
    L.panib_sketch_ascii_host.argtypes = [_vp, _vp, _i64, *sk_args, _vp, _vp, _vp, _vp]
    L.panib_sketch_ascii_host_hash_only.restype = _i32
    L.panib_sketch_ascii_host_hash_only.argtypes = [_vp, _vp, _i64, *sk_args, _vp, _vp, _vp]
    L.panib_set_workspace.restype = _i32
    L.panib_set_workspace.argtypes = [_vp, _i64]
    L.panib_host_threads.restype = _i32
    L.panib_pack_host.restype = _i32
    L.panib_pack_host.argtypes = [_vp, _i64, _vp, _vp, _i32]
    L.panib_pack_host_tiles.restype = _i32
    L.panib_pack_host_tiles.argtypes = [_vp, _i64, _vp, _vp, _vp, _i32]
    L.panib_sketch_packed_host.restype = _i32
    L.panib_sketch_packed_host.argtypes = [_vp, _vp, _vp, _i64, _vp, _i64, _vp, _vp, _vp, _i64, _i64, _i32, _u32, _u64,
                                           _vp, _vp, _vp, _i64, _vp, _vp, _vp, _i32, _vp]
    L.panib_ingest_scratch_bytes.restype = _i64
    L.panib_ingest_scratch_bytes.argtypes = [_i64]
    L.panib_ingest_last.restype = _i32
    L.panib_ingest_last.argtypes = [_vp]
    L.panib_sketch_finalize_gather.restype = _i32
    L.panib_sketch_finalize_gather.argtypes = [_vp, _i64, _i64, _vp, _vp, _vp, _vp, ctypes.POINTER(_vp), _i32, _i32,
                                               _i64, _vp]
    L.panib_intersect.restype = _i32
    L.panib_intersect.argtypes = [_vp, _vp, _i64, _i64, _vp, _vp, _i64, _i64, _i32, _u64, _i64, _i32, _i32,
                                  _i32, _vp, _vp, _i64, _i32, _i32, _vp, _vp]
    L.panib_intersect_fence_entries.restype = _i64
    L.panib_intersect_fence_entries.argtypes = [_i64, _i64, _u64, _i64, _i32, _i32]
    L.panib_index_workspace_bytes.restype = _i32
    L.panib_index_workspace_bytes.argtypes = [_i64, _i64, _i32, _i32, ctypes.POINTER(_i64)]
    L.panib_index_build.restype = _i32
    L.panib_index_build.argtypes = [_vp, _vp, _i64, _i64, _u64, _i64, _i64, _vp, _i32, _i32, _i32, _vp, _i64, _vp,
                                    _vp, _vp]
    L.panib_index_count.restype = _i32
    L.panib_index_count.argtypes = [_vp, _i64, _u64, _i64, _i64, _vp, _i32, _vp, _i64, _vp, _vp, _i64, _i32, _i32,
                                    _vp, _vp]
    L.panib_ani_device.restype = _i32
    L.panib_ani_device.argtypes = [_vp, _i64, _vp, _i64, _vp, _i64, _i32, _vp, _vp, _vp]
    L.panib_ani_host.restype = _i32
    L.panib_ani_host.argtypes = [_vp, _i64, _vp, _i64, _vp, _i64, _i32, _vp, _vp]
    L.panib_fasta_to_stream.restype = _i64
    L.panib_fasta_to_stream.argtypes = [ctypes.c_char_p, _i64, _vp, _i64, ctypes.POINTER(_i64)]
    L.panib_format_u64.restype = _i64
    L.panib_format_u64.argtypes = [_vp, _i64, _i32, _vp, _i64]
    L.panib_synth_ascii.restype = _i32
    L.panib_synth_ascii.argtypes = [_u64, _i64, _i64, _i64, _vp, _vp]
    _lib = L
    return L


def library_version() -> str:
    return load_library().panib_version().decode()


def max_hash(scaled: int) -> int:
    """sourmash's max_hash for a ``scaled`` value (host arithmetic inside the C ABI library)."""
    return int(load_library().panib_max_hash(scaled))


def plan_buckets(n_kmers: int, scaled: int, slack: float = 1.0) -> tuple[int, int]:
    nb = ctypes.c_int32()
    bmul = _u64()
    _check(load_library().panib_plan_buckets(n_kmers, scaled, slack, ctypes.byref(nb), ctypes.byref(bmul)))
    return int(nb.value), int(bmul.value)


def _check(rc: int) -> None:
    if rc != 0:
        buf = ctypes.create_string_buffer(512)
        load_library().panib_last_error(buf, 512)
        msg = f"libpanib200 error {rc}: {buf.value.decode(errors='replace')}"
        raise EngineError(msg)


def fasta_to_stream(text: bytes) -> tuple[np.ndarray, int, int, bytes | None]:
    """Parse decompressed FASTA text in C (``panib_fasta_to_stream``).

    Returns (stream-form uint8 array: records back to back with one ``N`` between them, number of
    records, total bases, title of the first record or None).  Same record semantics as
    ``utils.fasta_bytes_iterator`` / the reference's ``pyani_plus/utils.py:40-90``.
    """
    lib = load_library()
    out4 = (_i64 * 4)()
    # one pass: the stream form is never longer than the text (every record gives up its title line and gets at
    # most one separator), so the text's size is a safe destination and the result is a view of its front
    buf = np.empty(max(len(text), 1), dtype=np.uint8)
    got = int(lib.panib_fasta_to_stream(text, len(text), buf.ctypes.data, buf.size, out4))
    if got < 0:
        _check(got)
    title = text[out4[2]: out4[2] + out4[3]] if out4[2] >= 0 else None
    return buf[:got], int(out4[0]), int(out4[1]), title


def format_u64(values: np.ndarray, sep: bytes = b",") -> bytes:
    """Decimal text of uint64 values, ``sep`` (one byte, or b"" for none) between them (``panib_format_u64``)."""
    v = np.ascontiguousarray(values, dtype=np.uint64)
    if len(sep) > 1:
        msg = "sep must be one byte or empty"
        raise ValueError(msg)
    out = ctypes.create_string_buffer(21 * v.size + 1)
    got = int(load_library().panib_format_u64(v.ctypes.data, v.size, sep[0] if sep else 0, out, 21 * v.size + 1))
    if got < 0:
        _check(got)
    return ctypes.string_at(out, got)


def pack_host(ascii_stream: np.ndarray, threads: int = 0) -> tuple[np.ndarray, np.ndarray]:
    """ASCII base stream -> (packed 2-bit words, validity-mask words) on the library's host threads
    (``panib_pack_host``: the host form of the device pack kernel; needs no GPU)."""
    a = np.ascontiguousarray(ascii_stream, dtype=np.uint8)
    if a.size % 32:
        msg = "the ASCII stream length must be a multiple of 32"
        raise ValueError(msg)
    packed = np.empty(a.size // 16, dtype=np.uint32)
    mask = np.empty(a.size // 32, dtype=np.uint32)
    _check(load_library().panib_pack_host(a.ctypes.data, a.size, packed.ctypes.data, mask.ctypes.data, threads))
    return packed, mask


def ani_host(ov: np.ndarray, q_counts: np.ndarray, s_counts: np.ndarray, k: int) -> tuple[np.ndarray, np.ndarray]:
    """(identity, cov_query) float64 matrices from intersection counts, libm ``pow`` on the host.

    NaN where branchwater would print no row.  pyani-plus mapping (private_cli.py:1875-1887):
    identity = max_containment_ani, cov_query = query_containment_ani.
    """
    ov = np.ascontiguousarray(ov, dtype=np.uint32)
    q_counts = np.ascontiguousarray(q_counts, dtype=np.int32)
    s_counts = np.ascontiguousarray(s_counts, dtype=np.int32)
    nq, ns = ov.shape
    ident = np.empty((nq, ns), dtype=np.float64)
    cov = np.empty((nq, ns), dtype=np.float64)
    _check(load_library().panib_ani_host(ov.ctypes.data, ns, q_counts.ctypes.data, nq, s_counts.ctypes.data, ns,
                                         k, ident.ctypes.data, cov.ctypes.data))
    return ident, cov


@dataclass
class SketchTable:
    """Device-resident sketches: row g holds counts[g] ascending distinct hashes."""

    rows: "torch.Tensor"  # noqa: F821, UP037  int64 storage of uint64 hashes, shape [n, stride]
    counts: "torch.Tensor"  # noqa: F821, UP037  int32 [n]
    k: int
    scaled: int

    @property
    def n(self) -> int:
        return int(self.rows.shape[0])

    @property
    def stride(self) -> int:
        return int(self.rows.shape[1])

    def to_host(self) -> list[np.ndarray]:
        counts = self.counts.cpu().numpy()
        rows = self.rows.cpu().numpy().view(np.uint64)
        return [rows[g, : counts[g]].copy() for g in range(self.n)]


@dataclass
class StreamPlan:
    """Geometry + bucket plan of one tiled base stream (host and device copies)."""

    tile_off: np.ndarray
    n_genomes: int
    n_tiles: int
    n_bases: int
    scaled: int
    max_hash: int
    slack: float
    row_stride: int
    d_tile_off: "torch.Tensor"  # noqa: F821, UP037
    d_nb: "torch.Tensor"  # noqa: F821, UP037
    d_bmul: "torch.Tensor"  # noqa: F821, UP037

    def sketch_size_hint(self) -> int:
        """Upper estimate of the largest sketch of this stream: expected survivors of the longest
        genome + 6 sigma.  Lets K2 size its shared memory without a device->host read; the kernel
        still verifies it (PANIB_ST_SEGMENT_OVERFLOW -> re-plan)."""
        tiles = int(np.max(np.diff(self.tile_off))) if self.n_genomes else 1
        expected = tiles * _stream.TILE / max(1, self.scaled)
        return int(expected + 6.0 * expected ** 0.5 + 64)


class Engine:
    """One engine per process / per GPU (``torch.cuda.current_device()``)."""

    def __init__(self, device: int | None = None) -> None:
        import torch  # noqa: PLC0415

        self.torch = torch
        self.lib = load_library()
        if not torch.cuda.is_available() or self.lib.panib_device_count() < 1:
            msg = "No CUDA device visible: the sourmash path of pyani_plus_b200 has no CPU fallback."
            raise EngineError(msg)
        if device is None:
            device = torch.cuda.current_device()
        self.device = torch.device("cuda", device)
        torch.cuda.set_device(self.device)
        self.status = torch.zeros(4, dtype=torch.int32, device=self.device)
        self.last_max_count = 0
        self.last_intersect_method = ""
        self._index_work = None            # inverted-index K2: scratch of the checked (eager) calls
        self._index_work_fixed: dict = {}  # ... and of enqueue-only calls, by (n, entries, tau): graph-safe
        self._index_stats = torch.zeros(4, dtype=torch.int64, device=self.device)
        self.last_intersect_estimates: dict = {}
        self.max_batch_bytes = 1 << 31  # ASCII bytes staged per sketch batch

    # ------------------------------------------------------------------ helpers
    def _stream(self) -> int:
        return int(self.torch.cuda.current_stream(self.device).cuda_stream)

    def launch_count(self) -> int:
        return int(self.lib.panib_launch_count())

    def _read_status(self) -> int:
        """One device->host read (synchronises): status bits; remembers the largest sketch size seen."""
        st, max_count = self.status[:2].tolist()
        self.last_max_count = int(max_count)
        if st or max_count:
            self.status.zero_()
        return int(st)

    # ------------------------------------------------------------------ stage 0+1: sketch
    def plan_stream(self, tile_off: np.ndarray, scaled: int, slack: float = 1.0, *,
                    row_stride: int | None = None) -> "StreamPlan":
        """Bucket plan + device copies of the per-genome arrays for one tiled base stream.

        ``row_stride`` forces the table's row stride (a multiple of 1024 slots, at least what this
        stream needs): ranks whose slices hold genomes of different lengths must agree on ONE stride
        before their rows are exchanged (``run.agree_row_stride``)."""
        torch = self.torch
        tile_off = np.ascontiguousarray(tile_off, dtype=np.int64)
        n_genomes = len(tile_off) - 1
        nb = np.ones(max(n_genomes, 1), dtype=np.int32)
        bmul = np.zeros(max(n_genomes, 1), dtype=np.uint64)
        cache: dict[int, tuple[int, int]] = {}
        for g in range(n_genomes):
            tiles = int(tile_off[g + 1] - tile_off[g])
            if tiles not in cache:
                cache[tiles] = plan_buckets(tiles * _stream.TILE, scaled, slack)
            nb[g], bmul[g] = cache[tiles]
        return StreamPlan(
            tile_off=tile_off, n_genomes=n_genomes, n_tiles=int(tile_off[-1]),
            n_bases=(int(tile_off[-1]) + 1) * _stream.TILE, scaled=scaled, max_hash=max_hash(scaled),
            slack=slack, row_stride=self._checked_stride(int(nb.max()) * 1024, row_stride),
            d_tile_off=torch.from_numpy(tile_off).to(self.device),
            d_nb=torch.from_numpy(nb).to(self.device),
            d_bmul=torch.from_numpy(bmul.view(np.int64)).to(self.device),
        )

    @staticmethod
    def _checked_stride(needed: int, forced: int | None) -> int:
        if forced is None:
            return needed
        if forced < needed or forced % 1024:
            msg = f"row_stride={forced} must be a multiple of 1024 and at least {needed}"
            raise ValueError(msg)
        return int(forced)

    def alloc_stream_buffers(self, plan: "StreamPlan", *, ascii_too: bool = False, host_packed: bool = False) -> dict:
        """Device buffers of the packed stream; ``ascii_too`` adds a device ASCII buffer (device-side pack),
        ``host_packed`` the pinned host buffers the ingest pipeline packs into (``sketch_host``)."""
        torch = self.torch
        bufs = {
            "packed": torch.empty(plan.n_bases // 16, dtype=torch.int32, device=self.device),
            "mask": torch.empty(plan.n_bases // 32, dtype=torch.int32, device=self.device),
        }
        if ascii_too:
            bufs["ascii"] = torch.empty(plan.n_bases, dtype=torch.uint8, device=self.device)
        if host_packed:
            bufs["h_packed"] = torch.empty(plan.n_bases // 16, dtype=torch.int32, pin_memory=True)
            bufs["h_mask"] = torch.empty(plan.n_bases // 32, dtype=torch.int32, pin_memory=True)
            self.add_ingest_scratch(plan, bufs)
        return bufs

    def add_ingest_scratch(self, plan: "StreamPlan", bufs: dict) -> None:
        """Device scratch of the ingest pipeline (``panib_ingest_scratch_bytes``): staging of the sparse
        validity mask and of the chunks that cross PCIe as plain ASCII."""
        need = int(self.lib.panib_ingest_scratch_bytes(plan.n_bases))
        bufs["ingest_scratch"] = self.torch.empty(max(need, 1), dtype=self.torch.uint8, device=self.device)

    def _ensure_workspace(self, plan: "StreamPlan") -> None:
        """Survivor scratch of the sketch kernels on this device (``panib_set_workspace``): one per process
        and device, shared by every ``Engine`` there, grown to what the largest stream planned so far needs
        and never freed.  Once a CUDA graph has captured kernels that point at it, it is no longer replaced
        (a too-small one only means K1 inserts directly, the results are the same)."""
        if os.environ.get("PANIB_NO_WORKSPACE"):  # A/B knob: K1 then inserts survivors directly
            return
        need = int(2 * 12 * plan.n_bases / max(1, plan.scaled)) + (1 << 20)
        key = self.device.index
        have, frozen = _WORKSPACES.get(key, (None, False))
        if (have is not None and have.numel() >= need) or frozen:
            return
        ws = self.torch.empty(need, dtype=self.torch.uint8, device=self.device)
        _check(self.lib.panib_set_workspace(ws.data_ptr(), need))
        _WORKSPACES[key] = (ws, False)

    def freeze_workspace(self) -> None:
        """Called before a CUDA graph captures sketch kernels: the scratch they point at must stay."""
        key = self.device.index
        if key in _WORKSPACES:
            _WORKSPACES[key] = (_WORKSPACES[key][0], True)

    def alloc_table(self, plan: "StreamPlan") -> dict:
        torch = self.torch
        self._ensure_workspace(plan)
        n = max(plan.n_genomes, 1)
        return {
            "table": torch.empty((n, plan.row_stride), dtype=torch.int64, device=self.device),
            "counts": torch.zeros(n, dtype=torch.int32, device=self.device),
            "flags": torch.empty(n, dtype=torch.int32, device=self.device),
        }

    def pack(self, d_ascii, plan: "StreamPlan", bufs: dict) -> None:
        """ASCII device stream -> packed 2-bit + mask (stage 0)."""
        _check(self.lib.panib_pack_ascii(d_ascii.data_ptr(), plan.n_bases, bufs["packed"].data_ptr(),
                                         bufs["mask"].data_ptr(), self._stream()))

    def _sketch_args(self, plan: "StreamPlan", bufs: dict, tab: dict, k: int, seed: int) -> tuple:
        return (bufs["packed"].data_ptr(), bufs["mask"].data_ptr(), plan.d_tile_off.data_ptr(), plan.n_genomes,
                plan.n_tiles, k, seed, plan.max_hash, plan.d_nb.data_ptr(), plan.d_bmul.data_ptr(),
                tab["table"].data_ptr(), plan.row_stride, tab["counts"].data_ptr(), tab["flags"].data_ptr(),
                self.status.data_ptr(), self._stream())

    def sketch_packed(self, plan: "StreamPlan", bufs: dict, tab: dict, k: int, *, seed: int = 42) -> None:
        """Kernel K1 on a device-resident packed stream (enqueue only; see ``check_status``)."""
        _check(self.lib.panib_sketch_stream(*self._sketch_args(plan, bufs, tab, k, seed)))

    def sketch_ascii_host(self, h_ascii, plan: "StreamPlan", bufs: dict, tab: dict, k: int, *,
                          seed: int = 42) -> None:
        """H2D copy of a pinned ASCII stream + pack + K1 in one C-ABI call (enqueue only)."""
        _check(self.lib.panib_sketch_ascii_host(h_ascii.data_ptr(), bufs["ascii"].data_ptr(), plan.n_bases,
                                                *self._sketch_args(plan, bufs, tab, k, seed)))

    def sketch_host(self, h_ascii, plan: "StreamPlan", bufs: dict, tab: dict, k: int, *, seed: int = 42,
                    finalize: bool = True, threads: int = 0) -> None:
        """The ingest pipeline in one C-ABI call (``panib_sketch_packed_host``): the host threads pack the
        ASCII stream (a host tensor; ``None`` = ``bufs["h_packed"]`` / ``["h_mask"]`` are already filled)
        chunk by chunk into the pinned buffers while earlier chunks are copied (0.25 byte per base plus the
        masks of the tiles that hold invalid bases) and hashed; with ``bufs["ingest_scratch"]`` and a pinned
        ``h_ascii`` chunks also travel as ASCII from the tail of the stream whenever the link would idle.  Returns when the host work is done and the device work is enqueued; ``finalize=False``
        leaves the rows bucketed (multi-GPU: ``finalize_gather`` follows)."""
        a = self._sketch_args(plan, bufs, tab, k, seed)
        scratch = bufs.get("ingest_scratch")
        _check(self.lib.panib_sketch_packed_host(
            h_ascii.data_ptr() if h_ascii is not None else None, bufs["h_packed"].data_ptr(),
            bufs["h_mask"].data_ptr(), plan.n_bases, scratch.data_ptr() if scratch is not None else None,
            scratch.numel() if scratch is not None else 0, *a[:12], a[12] if finalize else None, a[13], a[14],
            threads, a[15]))

    def ingest_last(self) -> dict:
        """What the last ``sketch_host`` call of this thread moved over PCIe (``panib_ingest_last``)."""
        out = (ctypes.c_int64 * 6)()
        _check(self.lib.panib_ingest_last(out))
        return {"h2d_bytes": int(out[0]), "chunks": int(out[1]), "chunks_as_ascii": int(out[2]),
                "dirty_tiles": int(out[3]), "ring_bytes": int(out[4]), "measuring_call": bool(out[5])}

    def hash_packed(self, plan: "StreamPlan", bufs: dict, tab: dict, k: int, *, seed: int = 42) -> None:
        """K1 hashing only: rows are left as bucketed hash sets (finalize separately)."""
        a = self._sketch_args(plan, bufs, tab, k, seed)
        _check(self.lib.panib_sketch_hash_only(*a[:12], a[13], a[14], a[15]))

    def hash_ascii_host(self, h_ascii, plan: "StreamPlan", bufs: dict, tab: dict, k: int, *, seed: int = 42) -> None:
        """H2D + pack + K1 hashing only (finalize separately)."""
        a = self._sketch_args(plan, bufs, tab, k, seed)
        _check(self.lib.panib_sketch_ascii_host_hash_only(h_ascii.data_ptr(), bufs["ascii"].data_ptr(),
                                                          plan.n_bases, *a[:12], a[13], a[14], a[15]))

    def finalize(self, plan: "StreamPlan", tab: dict) -> None:
        """Sort / dedup / compact the rows in place (single-GPU finalize)."""
        _check(self.lib.panib_sketch_finalize(tab["table"].data_ptr(), plan.row_stride, plan.n_genomes,
                                              plan.d_nb.data_ptr(), tab["counts"].data_ptr(),
                                              tab["flags"].data_ptr(), self.status.data_ptr(), self._stream()))

    def finalize_gather(self, plan: "StreamPlan", tab: dict, peer_ptrs: list[int], rank: int, per_rank: int) -> None:
        """Finalize fused with the all-gather: sorted sketches are written into every rank's gathered
        table through the peer pointers (``multi_gpu.SymmetricGather``)."""
        arr = (_vp * len(peer_ptrs))(*peer_ptrs)
        _check(self.lib.panib_sketch_finalize_gather(
            tab["table"].data_ptr(), plan.row_stride, plan.n_genomes, plan.d_nb.data_ptr(),
            tab["counts"].data_ptr(), tab["flags"].data_ptr(), self.status.data_ptr(), arr, len(peer_ptrs), rank,
            per_rank, self._stream()))

    def check_status(self) -> int:
        """Synchronise and return (then clear) the PANIB_ST_* bits kernels raised."""
        return self._read_status()

    def sketch_ascii_stream(
        self, ascii_stream, tile_off: np.ndarray, k: int, scaled: int, *, seed: int = 42, slack: float = 1.0,
        from_host: bool = True,
    ) -> SketchTable:
        """Sketch every genome of one tiled ASCII base stream (``stream.py`` layout).

        ``ascii_stream`` is a host uint8 torch tensor when ``from_host`` (packed on the host threads and
        copied H2D in packed form inside the C call), else a device tensor.  Retries with more buckets if
        a bucket overflowed.
        """
        while True:
            plan = self.plan_stream(tile_off, scaled, slack)
            if ascii_stream.numel() != plan.n_bases:
                msg = f"ASCII stream has {ascii_stream.numel()} bytes, expected {plan.n_bases}"
                raise ValueError(msg)
            bufs = self.alloc_stream_buffers(plan, host_packed=from_host)
            tab = self.alloc_table(plan)
            if from_host:
                self.sketch_host(ascii_stream, plan, bufs, tab, k, seed=seed)
            else:
                self.pack(ascii_stream, plan, bufs)
                self.sketch_packed(plan, bufs, tab, k, seed=seed)
            st = self.check_status()
            if st & ST_BUCKET_OVERFLOW:
                if slack > 64:
                    msg = "sketch buckets keep overflowing (more than 64x the expected number of hashes)"
                    raise EngineError(msg)
                slack *= 2
                continue
            return SketchTable(tab["table"][: plan.n_genomes], tab["counts"][: plan.n_genomes], k, scaled)

    def sketch_genomes(self, genomes: list, k: int, scaled: int, *, seed: int = 42) -> SketchTable:
        """Sketch genomes given as lists of record sequences (what a FASTA parser yields) or as
        stream-form uint8 arrays (records already joined by one ``N``, see ``fasta_to_stream``).

        Genomes are staged in batches of at most ``max_batch_bytes`` of pinned host memory; the
        per-batch tables are compacted into one table whose stride is the largest sketch size.
        """
        torch = self.torch
        genomes = [g if isinstance(g, np.ndarray) else np.frombuffer(b"N".join(g), dtype=np.uint8) for g in genomes]
        lengths = [int(g.size) for g in genomes]
        batches: list[tuple[int, int]] = []
        start, acc = 0, 0
        for i, length in enumerate(lengths):
            need = (length // _stream.TILE + 1) * _stream.TILE
            if i > start and acc + need > self.max_batch_bytes:
                batches.append((start, i))
                start, acc = i, 0
            acc += need
        if genomes:
            batches.append((start, len(genomes)))
        parts: list[SketchTable] = []
        for b0, b1 in batches:
            tile_off = _stream.plan_tiles(lengths[b0:b1])
            nbytes = _stream.stream_bytes(tile_off)
            h_ascii = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
            _stream.fill_ascii_stream(h_ascii.numpy(), tile_off, [[g] for g in genomes[b0:b1]])
            parts.append(self.sketch_ascii_stream(h_ascii, tile_off, k, scaled, seed=seed))
        return self.concat_tables(parts, k, scaled)

    def concat_tables(self, parts: list[SketchTable], k: int, scaled: int) -> SketchTable:
        """One table with a tight stride (largest sketch, rounded up to 16 slots = 128 bytes)."""
        torch = self.torch
        if not parts:
            return SketchTable(torch.empty((0, 16), dtype=torch.int64, device=self.device),
                               torch.empty(0, dtype=torch.int32, device=self.device), k, scaled)
        max_count = max(int(p.counts.max().item()) if p.n else 0 for p in parts)
        stride = max(16, (max_count + 15) // 16 * 16)
        n = sum(p.n for p in parts)
        rows = torch.empty((n, stride), dtype=torch.int64, device=self.device)
        at = 0
        for p in parts:
            w = min(stride, p.stride)
            rows[at: at + p.n, :w] = p.rows[:, :w]
            at += p.n
        counts = torch.cat([p.counts for p in parts])
        return SketchTable(rows, counts, k, scaled)

    def table_from_host(self, sketches: list[np.ndarray], k: int, scaled: int, *, stride: int | None = None,
                        rows: int | None = None) -> SketchTable:
        """Upload sorted duplicate-free uint64 sketches (e.g. read back from ``.sig`` files).

        The kernels rely on every row being strictly ascending and <= max_hash (fence binary search,
        monotone bucket index, hash-range sharding), and cache files can come from anywhere, so that is
        checked here.  ``stride`` / ``rows`` pad the table (multi-GPU: all ranks exchange one shape)."""
        torch = self.torch
        n = len(sketches) if rows is None else int(rows)
        max_count = max((len(s) for s in sketches), default=0)
        need = max(16, (max_count + 15) // 16 * 16)
        if stride is None:
            stride = need
        if stride < need or n < len(sketches):
            msg = f"table_from_host: stride={stride} / rows={n} too small for {len(sketches)} sketches of up to {max_count}"
            raise ValueError(msg)
        mh = max_hash(scaled)
        host = np.zeros((n, stride), dtype=np.uint64)
        counts = np.zeros(n, dtype=np.int32)
        for g, s in enumerate(sketches):
            s = np.asarray(s, dtype=np.uint64)
            if s.size and (not (s[1:] > s[:-1]).all() or int(s[-1]) > mh):
                msg = f"sketch {g} is not strictly ascending within 0..max_hash (scaled={scaled}): corrupt cache file?"
                raise ValueError(msg)
            host[g, : len(s)] = s
            counts[g] = len(s)
        return SketchTable(torch.from_numpy(host.view(np.int64)).to(self.device),
                           torch.from_numpy(counts).to(self.device), k, scaled)

    # ------------------------------------------------------------------ stage 2: intersect
    def intersect(self, q: SketchTable, s: SketchTable | None = None, *, rank: int = 0, world: int = 1,
                  n_cells: int = 0, seg_cap: int = 0, idx_buckets: int = 0, max_count: int | None = None,
                  check: bool = True, method: str = "auto", tau: int = 0):
        """uint32 intersection sizes as an int32 torch tensor [nq, ns] (device).

        ``s is None`` = all-vs-all within ``q`` (only q < s is computed, the result is mirrored and
        the diagonal holds the sketch sizes).  ``check=False`` only enqueues (no device->host read, so
        the call can be captured into a CUDA graph; needs ``max_count``): the caller reads
        ``check_status()`` afterwards and re-plans itself on ``ST_SEGMENT_OVERFLOW``.

        ``method``: ``"probe"`` = the shared-memory probing kernel (cost ~ pairs x sketch size, any
        shape), ``"index"`` = the inverted-index form (sort by hash, bit matrix for frequent hashes,
        pair expansion for rare ones; cost ~ what the genomes share; all-vs-all only), ``"auto"`` =
        build the index when the probing estimate is not negligible and take the cheaper of the two
        from the index statistics (one small device->host read; with ``check=False`` auto means
        probe).  The counts are identical; ``last_intersect_method`` tells which one ran.
        """
        if method not in ("auto", "probe", "index"):
            msg = f"unknown intersect method {method!r}"
            raise ValueError(msg)
        torch = self.torch
        symmetric = s is None
        if symmetric:
            s = q
        if (q.k, q.scaled) != (s.k, s.scaled):
            msg = "query and subject sketches use different k / scaled"
            raise ValueError(msg)
        nq, ns = q.n, s.n
        ov = torch.empty((nq, ns), dtype=torch.int32, device=self.device)
        if nq == 0 or ns == 0:
            return ov
        mh = max_hash(q.scaled)
        if max_count is None:
            if not check:
                msg = "intersect(check=False) needs max_count (no device read is allowed)"
                raise ValueError(msg)
            max_count = int(torch.maximum(q.counts.max(), s.counts.max()).item())
        indexable = symmetric and nq >= 2 and mh < (1 << 64) - 16
        if method == "index" and not indexable:
            msg = "intersect(method='index') needs an all-vs-all call and scaled >= 2"
            raise ValueError(msg)
        if method == "auto" and (not indexable or not check):
            method = "probe"
        if method != "probe":
            done = self._intersect_indexed(q, ov, mh, max_count, tau, rank, world, check, method == "auto")
            if done is not None:
                return done
        self.last_intersect_method = "probe"
        cells = n_cells
        while True:
            nfence = int(self.lib.panib_intersect_fence_entries(nq, ns, mh, max_count, cells, seg_cap))
            fence = torch.empty(max(nfence, 1), dtype=torch.int32, device=self.device)
            _check(self.lib.panib_intersect(
                q.rows.data_ptr(), q.counts.data_ptr(), q.stride, nq,
                s.rows.data_ptr(), s.counts.data_ptr(), s.stride, ns,
                1 if symmetric else 0, mh, max_count, cells, seg_cap, idx_buckets,
                fence.data_ptr(), ov.data_ptr(), ns, rank, world, self.status.data_ptr(), self._stream()))
            if not check:
                return ov
            st = self._read_status()
            if st & ST_BUCKET_OVERFLOW:
                msg = "a sketch bucket overflowed in an earlier sketch call whose status was not checked"
                raise EngineError(msg)
            if st & ST_SEGMENT_OVERFLOW:
                cells = max(2, cells * 2) if cells else max(2, 2 * -(-max_count // 4096))
                if cells > 1 << 16:
                    msg = "pairwise segments keep overflowing"
                    raise EngineError(msg)
                continue
            return ov

    # cost model of Engine.intersect(method="auto"), seconds per unit on one B200 (measured)
    COST_PROBE_PER_ELEMENT = 0.6e-12   # probing kernel: per pair and per sketch element (both sketches)
    COST_INDEX_PER_ENTRY = 0.6e-10     # hash-table insert + classify + emit, per (hash, genome) entry
    COST_INDEX_PER_WORD = 0.8e-12      # AND+POPC kernel: per pair and per 32 bit-matrix columns
    COST_INDEX_PER_RARE_PAIR = 2.0e-11 # atomicAdd expansion of the rare hashes, per pair

    def _intersect_indexed(self, q: SketchTable, ov, mh: int, max_count: int, tau: int, rank: int, world: int,  # noqa: ANN001, PLR0913
                           check: bool, auto: bool):  # noqa: ANN202
        """Inverted-index form of the all-vs-all intersection (csrc/index.cu).  Returns ``ov``, or None
        when ``auto`` decided that the probing kernel is cheaper for this data (or the index cannot be used)."""
        torch = self.torch
        n = q.n
        all_pairs = n * (n - 1) / 2
        est_probe = all_pairs / world * 2 * max_count * self.COST_PROBE_PER_ELEMENT
        if auto and est_probe < 3e-4:  # the index has ~8 launches of fixed cost: not worth it
            return None
        while True:
            cap = max(1, int(max_count))
            tau_ = tau if tau >= 2 else max(8, n // 32)
            if check:  # exact entries: prefix sums of the sizes (one small read for their total)
                offsets = torch.zeros(n + 1, dtype=torch.int64, device=self.device)
                offsets[1:] = torch.cumsum(q.counts, 0, dtype=torch.int64)
                total = int(offsets[-1].item())
                off_ptr = offsets.data_ptr()
                if total == 0 or total >= (1 << 31) - 512:
                    return None  # nothing to index / too many entries: the probing kernel handles both
            else:  # no device read allowed (graph capture): cap slots per genome, padded
                offsets, total, off_ptr = None, n * cap, None
                if total >= (1 << 31) - 512:
                    return None
            need = _i64(0)
            _check(self.lib.panib_index_workspace_bytes(n, total, tau_, world, ctypes.byref(need)))
            if check:  # eager calls share one workspace that grows on demand
                work = self._index_work
                if work is None or work.numel() < need.value:
                    self._index_work = work = None  # free the old one first
                    self._index_work = work = torch.empty(need.value, dtype=torch.uint8, device=self.device)
            else:  # enqueue-only calls may sit in a captured graph: their workspace is never freed or resized
                key = (n, total, tau_, world)
                work = self._index_work_fixed.get(key)
                if work is None:
                    work = self._index_work_fixed[key] = torch.empty(need.value, dtype=torch.uint8,
                                                                     device=self.device)
            stats = self._index_stats
            _check(self.lib.panib_index_build(q.rows.data_ptr(), q.counts.data_ptr(), q.stride, n, mh, cap, total,
                                              off_ptr, tau_, rank, world, work.data_ptr(), work.numel(),
                                              stats.data_ptr(), self.status.data_ptr(), self._stream()))
            if auto:
                n_dense, rare_pairs, _, _ = stats.tolist()  # of this rank's slice of the hash range
                est_index = (total / world * self.COST_INDEX_PER_ENTRY
                             + all_pairs * -(-n_dense // 32) * self.COST_INDEX_PER_WORD
                             + rare_pairs * self.COST_INDEX_PER_RARE_PAIR)
                self.last_intersect_estimates = {"probe_s": est_probe, "index_s": est_index,
                                                 "frequent_hashes": n_dense, "rare_pairs": rare_pairs}
                if est_index >= est_probe:
                    return None
            _check(self.lib.panib_index_count(q.counts.data_ptr(), n, mh, cap, total, off_ptr, tau_, work.data_ptr(),
                                              work.numel(), stats.data_ptr(), ov.data_ptr(), q.n, rank, world,
                                              self.status.data_ptr(), self._stream()))
            self.last_intersect_method = "index"
            if not check:
                return ov
            st = self._read_status()
            if st & ST_BUCKET_OVERFLOW:
                msg = "a sketch bucket overflowed in an earlier sketch call whose status was not checked"
                raise EngineError(msg)
            if st & ST_INDEX_OVERFLOW:
                msg = ("the hash range is too unevenly filled for the sharded inverted index "
                       "(one rank's slice holds more than twice its share): use method='probe'")
                raise EngineError(msg)
            if st & ST_SEGMENT_OVERFLOW:  # a sketch is larger than cap: take the real maximum and redo
                max_count = int(q.counts.max().item())
                if max_count <= cap:
                    msg = "index intersect keeps overflowing"
                    raise EngineError(msg)
                continue
            return ov

    # ------------------------------------------------------------------ stage 3: ANI
    def ani_device(self, ov, q: SketchTable, s: SketchTable | None = None):
        """(identity, cov_query) float64 device tensors, NaN where there is no row (CUDA pow)."""
        torch = self.torch
        if s is None:
            s = q
        ident = torch.empty((q.n, s.n), dtype=torch.float64, device=self.device)
        cov = torch.empty((q.n, s.n), dtype=torch.float64, device=self.device)
        _check(self.lib.panib_ani_device(ov.data_ptr(), s.n, q.counts.data_ptr(), q.n, s.counts.data_ptr(), s.n,
                                         q.k, ident.data_ptr(), cov.data_ptr(), self._stream()))
        return ident, cov

    # ------------------------------------------------------------------ synthetic input (bench / tests)
    def synth_ascii_stream(self, seed: int, g0: int, n_genomes: int, length: int):
        """Device ASCII stream of synthetic genomes + its tile offsets (SURVEY.md 8d generator)."""
        torch = self.torch
        tiles_per = length // _stream.TILE + 1
        tile_off = np.arange(n_genomes + 1, dtype=np.int64) * tiles_per
        n_bases = (int(tile_off[-1]) + 1) * _stream.TILE
        d_ascii = torch.full((n_bases,), _stream.PAD, dtype=torch.uint8, device=self.device)
        _check(self.lib.panib_synth_ascii(seed, g0, n_genomes, length, d_ascii.data_ptr(), self._stream()))
        return d_ascii, tile_off
