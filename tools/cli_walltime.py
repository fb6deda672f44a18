"""Wall time of the whole drop-in CLI path on synthetic FASTA files (host ingest + GPU + SQLite).

    python tools/cli_walltime.py [n_genomes=100] [length=5000000] [scaled=1000] [runs=cold,warm]

For the 10,000-genome drop-in check use shorter genomes with a proportionally smaller ``scaled``
(``10000 500000 100``): sketches, K2 and the 10^8 database rows are those of the 5 Mb / scaled=1000
configuration, only the FASTA bytes on disk are 10x fewer (5 GB instead of 50 GB).
"""
import logging
import sys
import tempfile
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from oracle import oracle  # noqa: E402  (only to WRITE the input files)
from pyani_plus_b200 import db_orm, public_cli, setup_logger  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100
length = int(sys.argv[2]) if len(sys.argv) > 2 else 5_000_000
scaled = int(sys.argv[3]) if len(sys.argv) > 3 else 1000
runs = (sys.argv[4] if len(sys.argv) > 4 else "cold,warm").split(",")
with tempfile.TemporaryDirectory() as tmp:
    tmp = Path(tmp)
    fasta = tmp / "fasta"
    fasta.mkdir()
    t0 = time.perf_counter()
    from concurrent.futures import ThreadPoolExecutor  # noqa: E402, PLC0415

    def write_one(g: int) -> None:
        seq = oracle.synth_genome(20261017, g, length)  # ctypes call: releases the GIL
        with (fasta / f"g{g:05d}.fna").open("wb") as fh:
            fh.write(b">g%d synthetic\n" % g)
            fh.write(b"\n".join(seq[i:i + 80] for i in range(0, length, 80)))
            fh.write(b"\n")

    with ThreadPoolExecutor(max_workers=16) as pool:
        list(pool.map(write_one, range(n)))
    t1 = time.perf_counter()
    print(f"wrote {n} FASTA files of {length} bp in {t1 - t0:.1f} s")
    logging.disable(logging.INFO)
    t = time.perf_counter()
    from pyani_plus_b200.methods import sourmash as _sm  # noqa: E402, PLC0415

    _sm.get_engine()  # import torch + create the CUDA context (a minute on a freshly paged-in image)
    print(f"engine start-up (import torch, CUDA context): {time.perf_counter() - t:.1f} s (not part of the runs below)")
    for attempt in [a for a in ("cold", "warm-cache (.sig files exist)") if a[:4] in runs]:
        db = tmp / f"{attempt[:4]}.db"
        t2 = time.perf_counter()
        rc = public_cli.cli_sourmash(fasta=fasta, database=db, create_db=True, cache=tmp, scaled=scaled)
        t3 = time.perf_counter()
        with db_orm.connect_to_db(setup_logger(None), db) as session:
            (run,) = session.runs()
            rows = run.comparisons().count()
        print(f"{attempt}: pyani-plus sourmash on {n} genomes -> rc={rc}, {rows} comparisons, {t3 - t2:.2f} s wall, "
              f"database {db.stat().st_size / 1e6:.0f} MB", flush=True)
