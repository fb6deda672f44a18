"""Collect the reference's own golden vectors for the sourmash path into tests/golden/.

Run in the build container (where /root/reference is mounted):

    python tests/golden/make_golden.py

The reference's arithmetic for this path is done by the external sourmash / branchwater tools,
which are not installable here, so nothing is *generated* by running the reference: the vectors
are the files the reference's own tests pin that path with:

* input FASTA files                          tests/fixtures/{viral_example,bad_alignments,bacterial_example}/*, MIBY*.fasta
* sourmash signature files (*.sig)           tests/fixtures/<set>/intermediates/sourmash/   (tests/snakemake/test_sourmash_workflow.py:43-109)
* branchwater manysearch.csv                 same directories                              (tests/test_sourmash.py:297-326)
* matrices/sourmash_{identity,coverage}.tsv  tests/fixtures/<set>/matrices/                (tests/snakemake/__init__.py:83-166)
* plots/sourmash_*_scatter.tsv               tests/fixtures/viral_example/plots/           (tests/test_public_cli.py:1053-1062)

plus ``expected.json`` holding the literal values asserted in tests/test_coverage.py:162-174 and
tests/test_self_vs_self.py and SURVEY.md Appendix A's single-k-mer answers.
Only DATA files are copied; no reference source code is.
"""

from __future__ import annotations

import json
import shutil
from pathlib import Path

REF = Path("/root/reference/tests/fixtures")
OUT = Path(__file__).resolve().parent


def copy(src: Path, dst: Path) -> None:
    dst.parent.mkdir(parents=True, exist_ok=True)
    shutil.copyfile(src, dst)
    dst.chmod(0o644)


def main() -> None:
    for name in ("viral_example", "bad_alignments", "bacterial_example"):
        src = REF / name
        for f in sorted(src.iterdir()):
            if f.is_file() and (f.suffix in {".fas", ".fna", ".fasta", ".fa"} or f.name.endswith(".gz")):
                copy(f, OUT / name / f.name)
        for f in sorted((src / "intermediates" / "sourmash").iterdir()):
            copy(f, OUT / name / "intermediates" / "sourmash" / f.name)
        for f in sorted((src / "matrices").glob("sourmash_*.tsv")):
            copy(f, OUT / name / "matrices" / f.name)
    for f in sorted((REF / "viral_example" / "plots").glob("sourmash_*.tsv")):
        copy(f, OUT / "viral_example" / "plots" / f.name)
    for name in ("MIBY01000005.fasta", "MIBY01000011.fasta"):
        copy(REF / name, OUT / name)

    expected = {
        "_source": "literals asserted by the reference's tests; see make_golden.py docstring",
        "test_coverage_scaled50": {
            # tests/test_coverage.py:54-80,162-174
            "files": {
                "small.fasta": "MIBY01000005.fasta",
                "large.fasta": "MIBY01000011.fasta",
                "both.fasta": ["MIBY01000005.fasta", "MIBY01000011.fasta"],
            },
            "md5_sorted": [
                "154173fb8e7415ab45532a738572f957",
                "7b6a6226ce00e52edca15565aa0d270d",
                "a0efc718e680e34d2f5c8f5d2286ca9c",
            ],
            "scaled": 50,
            "df_identity_data": [[1.0, 1.0, None], [1.0, 1.0, 1.0], [None, 1.0, 1.0]],
            "df_cov_query_data": [[1.0, 1.0, None], [0.9622440235, 1.0, 0.9884105907], [None, 1.0, 1.0]],
        },
        "single_kmers_k31_seed42": {
            # SURVEY.md Appendix A (derived from the restatement that reproduced every .sig golden)
            "AAAAAAAAAAAAAAAAAAAAAAAAAAAAAAA": 2689060699658636553,
            "TTTTTTTTTTTTTTTTTTTTTTTTTTTTTTT": 2689060699658636553,
            "ACGTACGTACGTACGTACGTACGTACGTACG": 17897553464741958189,
            "CCTTGTAGTTCAACGGATAGAATAGAAGTTT": 12833074019568967469,
        },
        "max_hash": {"300": 61489146912365176, "1000": 18446744073709552},
        "sketch_sizes": {
            # SURVEY.md Appendix A
            "689d3fd6881db36b5e08329cf23cecdd": 111,
            "78975d5144a1cd12e98898d573cf6536": 117,
            "5584c7029328dc48d33f95f0a78f7e57": 183,
            "a30481565b45f6bbc6ce5260503067e0": 270,
            "f19cb07198a41a4406a22b2f57a6b5e7": 4038,
            "073194224aa8c13bebc1d14a3e74a3e7": 5410,
            "9d72a8fb513cf9cc8cc6605a0ad4e837": 4067,
            "9a9e23bfc5a184b8149e07e267d133b0": 4495,
        },
    }
    (OUT / "expected.json").write_text(json.dumps(expected, indent=1) + "\n")
    print("golden vectors written under", OUT)


if __name__ == "__main__":
    main()
