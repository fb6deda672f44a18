"""The records of the path on stdlib sqlite3, exercised as the reference's tests/test_orm.py exercises its
SQLAlchemy models (test_make_and_populate_mock_example :269-476, test_add_config :479-524, test_insert_no_comps
:663-680): same reprs, same run-scoped comparison join, same cached matrices."""

from __future__ import annotations

import datetime
from pathlib import Path

import numpy as np
import pandas as pd

from pyani_plus_b200 import db_orm, setup_logger
from pyani_plus_b200.utils import str_md5sum


def test_make_and_populate_mock_example(tmp_path: Path) -> None:
    """Four genomes and all 16 comparisons recorded, but only two genomes linked to the run: the run sees 4
    comparisons and 2 x 2 matrices."""
    tmp_db = tmp_path / "mock.sqlite"
    logger = setup_logger(None)
    with db_orm.connect_to_db(logger, tmp_db) as session:
        config = db_orm.db_configuration(session, "guessing", "guestimate", "v0.1.2beta3", fragsize=1000, kmersize=31,
                                         create=True)
        assert repr(config) == ("Configuration(configuration_id=1, program='guestimate', version='v0.1.2beta3',"
                                " fragsize=1000, mode=None, kmersize=31, minmatch=None, extra=None)")
        when = datetime.datetime(2023, 12, 25)  # noqa: DTZ001
        cmdline = "pyani_plus run -m guestimate --input ../my-genomes/ -d working.sqlite"
        empty = db_orm.add_run(session, config, cmdline, Path("../my-genomes/"), "Aborted", "Empty", date=when)
        assert repr(empty) == (f"Run(run_id=1, configuration_id=1, cmdline={cmdline!r}, date={when!r},"
                               " status='Aborted', name='Empty', ...)")
        hashes, linked = [], {}
        for name in ("Genome A", "Genome C", "Genome G", "Genome T"):
            seq = (name[-1] + "ACGT") * 1000
            md5 = str_md5sum(f">{name}\n{seq}\n")
            hashes.append(md5)
            genome = db_orm.db_genome(logger, session, f"../my-genomes/{name}.fasta", md5, create=True,
                                      stats=(len(seq), name.encode(), False))
            assert repr(genome) == (f"Genome(genome_hash={md5!r}, path='../my-genomes/{name}.fasta',"
                                    f" length={len(seq)}, description='{name}')")
            if name[-1] in "AT":
                linked[Path(f"../my-genomes/{name}.fasta")] = md5
        run = db_orm.add_run(session, config, cmdline, Path("../my-genomes/"), "Complete", "Test Run", date=when,
                             fasta_to_hash=linked)
        uname = type("U", (), {"system": "Darwin", "release": "21.6.0", "machine": "arm64"})()
        for a in hashes:
            for b in hashes:
                comparison = db_orm.db_comparison(session, config.configuration_id, a, b, 0.99 if a == b else 0.96,
                                                  4996 if a == b else 4975, uname=uname)
                assert repr(comparison) == (
                    f"Comparison(comparison_id={comparison.comparison_id}, query_hash={a!r}, subject_hash={b!r},"
                    f" configuration_id=1, identity={0.99 if a == b else 0.96},"
                    f" aln_length={4996 if a == b else 4975}, sim_errors=None, cov_query=None, cov_subject=None,"
                    " uname_system='Darwin', uname_release='21.6.0', uname_machine='arm64')")
        assert run.identities is None  # not collated yet
        run.cache_comparisons()
        session.commit()
        two = ["0ac5f24c37ea8ef2d0e37fbfa61e2a43", "4bce437e7bdd91e35d18bfe294dee207"]
        assert sorted(linked.values()) == two

        def frame(values: list[list[float]]) -> pd.DataFrame:
            return pd.DataFrame(data=np.array(values, float), index=two, columns=two)

        nans = [[np.nan, np.nan], [np.nan, np.nan]]
        assert run.identities.equals(frame([[0.99, 0.96], [0.96, 0.99]]))
        assert run.cov_query.equals(frame(nans))
        assert run.aln_length.astype(float).equals(frame([[4996, 4975], [4975, 4996]]))
        assert run.sim_errors.equals(frame(nans))
        assert run.hadamard.equals(frame(nans))  # float * nan = nan

    with db_orm.connect_to_db(logger, tmp_db) as session:
        assert session.execute("SELECT COUNT(*) FROM genomes").fetchone()[0] == 4  # noqa: PLR2004
        assert session.execute("SELECT COUNT(*) FROM comparisons").fetchone()[0] == 16  # noqa: PLR2004
        (run,) = [r for r in session.runs() if r.status == "Complete"]
        assert run.configuration.configuration_id == 1 and run.configuration.method == "guessing"
        assert {g.genome_hash for g in run.genomes} == set(two)
        rows = list(run.comparisons())
        assert len(rows) == 4 and run.comparisons().count() == 4  # only 4, not all 16  # noqa: PLR2004
        assert {(c.query_hash, c.subject_hash) for c in rows} == {(q, s) for q in two for s in two}
        assert run.comparisons().null_count() == 0
        assert [r[:2] for r in run.comparisons().values()] == [(q, s) for q in two for s in two]
        assert run.identities.equals(pd.DataFrame(data=np.array([[0.99, 0.96], [0.96, 0.99]]), index=two, columns=two))


def test_add_config(tmp_path: Path) -> None:
    """db_configuration returns the existing row for identical settings and refuses to create when asked not to."""
    import pytest

    with db_orm.connect_to_db(setup_logger(None), tmp_path / "config.sqlite") as session:
        with pytest.raises(db_orm.NoResultFound, match="Requested configuration not already in DB"):
            db_orm.db_configuration(session, "guessing", "guestimate", "v0.1.2beta3", fragsize=100, kmersize=17,
                                    minmatch=0.3, create=False)
        first = db_orm.db_configuration(session, "guessing", "guestimate", "v0.1.2beta3", fragsize=100, kmersize=17,
                                        minmatch=0.3, create=True)
        again = db_orm.db_configuration(session, "guessing", "guestimate", "v0.1.2beta3", fragsize=100, kmersize=17,
                                        minmatch=0.3, create=False)
        other = db_orm.db_configuration(session, "guessing", "guestimate", "v0.1.2beta3", fragsize=100, kmersize=17,
                                        minmatch=0.3, extra="scaled=1000", create=True)
        assert (first.configuration_id, again.configuration_id, other.configuration_id) == (1, 1, 2)


def test_insert_no_comps(tmp_path: Path) -> None:
    """Recording nothing is a success (reference: insert_comparisons_with_retries on an empty list)."""
    logger = setup_logger(None)
    with db_orm.connect_to_db(logger, tmp_path / "empty.sqlite") as session:
        assert db_orm.insert_comparisons_with_retries(logger, session, [])
        config = db_orm.db_configuration(session, "sourmash", "panib200", "0", kmersize=31, extra="scaled=1000",
                                         create=True)
        assert db_orm.insert_comparison_arrays(logger, session, config.configuration_id, [], [], np.empty((0, 0)),
                                               np.empty((0, 0)))
        assert session.execute("SELECT COUNT(*) FROM comparisons").fetchone()[0] == 0
