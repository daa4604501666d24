"""Generates tests/golden/reference_host.json by running the REAL reference code in this container.

    python tests/golden/make_reference_golden.py        (needs /root/reference; the tests do not)

Imported from /root/reference, unmodified:
  * inference_pipeline/db_utils/setup_db.py  -- the sqlite helpers behind heavy_ranker.py:70-72,102-113.

Not importable here, so NOT pinned this way (their mirrors are checked by reading, tests/test_host.py):
  * src/data/configs/advance_qa_sample.py (straighten_docs :99-106) -> response_template.py:289 declares a
    dataclass field with a mutable default, which Python >= 3.11 rejects at import (ValueError), and
    ``src/data/__init__`` needs trl / peft;
  * everything behind ``import txtai`` (heavy_ranker.py:4).
"""
import importlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"


def load_reference():
    sys.path.insert(0, REF)                                         # setup_db does ``from src.utils import timeit``
    sys.path.insert(0, os.path.join(REF, "inference_pipeline", "db_utils"))
    cwd = os.getcwd()
    os.chdir(REF)
    try:
        setup_db = importlib.import_module("setup_db")
    finally:
        os.chdir(cwd)
    return setup_db


def main():
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    from tests.golden import host_scenario

    out = host_scenario.run(load_reference())
    with open(os.path.join(HERE, "reference_host.json"), "w", encoding="utf-8") as f:
        json.dump(out, f, ensure_ascii=False, indent=1, sort_keys=True)
    print(json.dumps(out, ensure_ascii=False)[:600])


if __name__ == "__main__":
    main()
