"""One scripted scenario over the sqlite helper API (setup_database / insert_data / query / drop_tables /
connect_database).  ``run(api)`` is executed

* by ``make_reference_golden.py`` against the REAL reference module imported from /root/reference
  (inference_pipeline/db_utils/setup_db.py) -> tests/golden/reference_host.json
* by ``tests/test_reference_golden.py`` against this repository's mirror (vietnamese_qa_system_b200.db),
  whose results must equal the committed reference outputs.

Exceptions are recorded by class name only: the reference's error paths do ``raise "<str>"`` (a TypeError in
Python 3), the mirror raises sqlite3.OperationalError / ValueError -- that one documented difference is
asserted separately by the test.
"""
import contextlib
import io
import os
import tempfile

ROWS = [{"doc": f"Đoạn văn số {i}: Hà Nội, phở, sông Hồng.", "source": f"wiki/{i % 3}"} for i in range(12)]


def _outcome(fn, *args, **kwargs):
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            return {"ok": _plain(fn(*args, **kwargs))}
    except BaseException as exc:  # noqa: BLE001 - the class is the recorded behaviour
        return {"raises": type(exc).__name__}


def _plain(x):
    if isinstance(x, tuple):
        return [_plain(v) for v in x]
    if isinstance(x, list):
        return [_plain(v) for v in x]
    return x


def run(api):
    """``api``: module-like object with the five helper functions."""
    out = {}
    with tempfile.TemporaryDirectory() as d:
        r = _outcome(api.setup_database, "documents", database_dir=d, verbose=True)
        path = os.path.join(d, "documents.db")
        out["setup_returns_path"] = r.get("ok") == path
        out["setup_twice"] = "ok" in _outcome(api.setup_database, "documents", database_dir=d, verbose=False)
        out["setup_bad_dir"] = _outcome(api.setup_database, "x", database_dir=os.path.join(d, "missing"))
        out["setup_len_mismatch"] = _outcome(api.setup_database, "y", table_names=["a", "b"], fields=["(id INTEGER)"],
                                             database_dir=d)
        out["insert"] = _outcome(api.insert_data, path, "documents", ROWS, verbose=True)
        out["insert_empty"] = _outcome(api.insert_data, path, "documents", [], verbose=False)
        out["insert_bad_table"] = _outcome(api.insert_data, path, "nope", ROWS[:1], verbose=False)
        out["insert_bad_column"] = _outcome(api.insert_data, path, "documents", [{"nope": 1}], verbose=False)
        q = "SELECT * FROM documents"
        out["query_all"] = _outcome(api.query, path, q)
        out["query_all_explicit"] = _outcome(api.query, path, q, fetch_size="all")
        out["query_many_5"] = _outcome(api.query, path, q, fetch_size=5, verbose=True)
        out["query_many_50000"] = _outcome(api.query, path, q, fetch_size=50000)       # heavy_ranker.py:70-72
        out["query_one"] = _outcome(api.query, path, "SELECT doc FROM documents WHERE id = 7", fetch_size=1)
        out["query_one_missing"] = _outcome(api.query, path, "SELECT doc FROM documents WHERE id = 700", fetch_size=1)
        out["query_where"] = _outcome(api.query, path, "SELECT id, source FROM documents WHERE source = 'wiki/1'")
        out["query_bad_sql"] = _outcome(api.query, path, "SELEKT nothing")
        out["query_bad_mode_zero"] = _outcome(api.query, path, q, fetch_size=0)
        out["query_bad_mode_str"] = _outcome(api.query, path, q, fetch_size="some")
        out["query_missing_file"] = _outcome(api.query, os.path.join(d, "absent.db"), q)
        txt = os.path.join(d, "notes.txt")
        with open(txt, "w") as f:
            f.write("x")
        out["connect_bad_extension"] = _outcome(api.connect_database, txt)
        out["connect_ok_type"] = type(api.connect_database(path)).__name__
        out["drop_missing_table"] = _outcome(api.drop_tables, path, ["nope"], verbose=False)
        out["drop"] = _outcome(api.drop_tables, path, ["documents"], verbose=True)
        out["query_after_drop"] = _outcome(api.query, path, q)
        out["recreate_ids_restart"] = None
        api_out = _outcome(api.setup_database, "documents", database_dir=d, verbose=False)
        if "ok" in api_out:
            _outcome(api.insert_data, path, "documents", ROWS[:2], verbose=False)
            out["recreate_ids_restart"] = _outcome(api.query, path, "SELECT id FROM documents")
    return out
