"""sqlite passage store -- mirror of the reference's helper module.

Same function names, argument order, defaults and return shapes as
``inference_pipeline/db_utils/setup_db.py`` (setup_database :12-37, drop_tables
:40-56, query :59-83, insert_data :86-116, connect_database :119-132), so that
``from vietnamese_qa_system_b200.db import setup_database, drop_tables, query,
insert_data`` replaces ``from setup_db import ...`` at heavy_ranker.py:6-7.

Differences, all on the error path: the reference does ``raise "<str>"`` (a
TypeError in Python 3); here the same conditions raise ``sqlite3.OperationalError``
/ ``ValueError`` with the reference's message.  ``fetch_docs`` is the batched
id -> passage lookup that follows a top-k: one
``SELECT ... WHERE id IN (...)`` instead of one connection per id
(heavy_ranker.py:102-109).
"""
from __future__ import annotations

import os
import sqlite3
from sqlite3 import Connection, OperationalError
from typing import Any, Dict, Iterable, List, Union


def connect_database(database_path: str, verbose: bool = False) -> Connection:
    assert os.path.isfile(database_path), f"Invalid database path for {database_path}"
    assert database_path[-2:] == "db" or database_path[-6:] == "sqlite", \
        "Invalid file, the file must have an extension .db or .sqlite"
    try:
        connection = sqlite3.connect(database_path)
    except OperationalError as e:
        raise OperationalError(f"Connection to database {database_path} failed with the following error\n"
                               f"Error message: {e}") from e
    if verbose:
        print(f"Connect to database {database_path} successfully")
    return connection


def setup_database(database_name: str,
                   table_names: List[str] = ["documents"],
                   fields: List[str] = ['''(id INTEGER PRIMARY KEY AUTOINCREMENT, doc TEXT, source TEXT)'''],
                   database_dir: str = "./inference_pipeline/dbs",
                   verbose: bool = True) -> str:
    assert os.path.isdir(database_dir), f"Invalid database_dir path: {database_dir}"
    assert len(table_names) == len(fields), "The table_names and the fields args must have the same length"
    database_path = os.path.join(database_dir, f"{database_name}.db")
    try:
        connection = sqlite3.connect(database_path)
    except OperationalError as e:
        raise OperationalError(f"Connection to database {database_name} failed with the following error\n"
                               f"Error message: {e}") from e
    if verbose:
        print(f"Successfully create database {database_path}")
    cursor = connection.cursor()
    try:
        for table_name, field in zip(table_names, fields):
            try:
                cursor.execute(f"CREATE TABLE IF NOT EXISTS {table_name} {field}")
            except OperationalError as e:
                raise OperationalError(f"Create table {table_name} fail with the following error: {e}") from e
            if verbose:
                print(f"Successfully create table {table_name} with field {field}")
        connection.commit()
    finally:
        connection.close()
    return database_path


def drop_tables(database_path: str, tables_to_drop: List[str], verbose: bool = True):
    connection = connect_database(database_path, verbose=verbose)
    cursor = connection.cursor()
    try:
        for table_name in tables_to_drop:
            try:
                cursor.execute(f"DROP TABLE {table_name}")
            except OperationalError as e:
                raise OperationalError(f"Cannot drop table {table_name} with the following error: {e}") from e
            if verbose:
                print(f"Successfully drop table {table_name}")
        connection.commit()
    finally:
        connection.close()
    if verbose:
        print(f"Drop tables: {tables_to_drop} successfully")


def query(database_path: str, query_string: str, fetch_size: Union[int, str] = "all",
          verbose: bool = False) -> Union[list, Any]:
    connection = connect_database(database_path, verbose=verbose)
    try:
        cursor = connection.cursor()
        try:
            cursor.execute(query_string)
        except OperationalError as e:
            raise OperationalError(f"Query {query_string} failed with the following error: {e}") from e
        if fetch_size == "all":
            if verbose:
                print("Fetch all rows")
            data = cursor.fetchall()
        elif isinstance(fetch_size, int) and fetch_size > 1:
            if verbose:
                print(f"Fetch {fetch_size} rows")
            data = cursor.fetchmany(size=fetch_size)
        elif fetch_size == 1:
            if verbose:
                print("Fetch 1 row")
            data = cursor.fetchone()
        else:
            raise ValueError("Invalid fetch mode")
    finally:
        connection.close()
    return data


def insert_data(database_path: str, table_name: str, data: List[dict], verbose: bool = True):
    connection = connect_database(database_path, verbose=verbose)
    cursor = connection.cursor()
    try:
        cursor.execute("BEGIN TRANSACTION")
        columns = ", ".join(data[0].keys()) if data else ""
        placeholders = ", ".join(["?"] * len(data[0])) if data else ""
        insert_query = f"INSERT INTO {table_name} ({columns}) VALUES ({placeholders})"
        if verbose:
            print(f"The query for insert: {insert_query}")
        values = [tuple(row.values()) for row in data]
        cursor.executemany(insert_query, values)
        connection.commit()
        if verbose:
            print(f"Successfully inserted {len(data)} rows into table {table_name} in {database_path}")
    except OperationalError as e:
        connection.rollback()
        raise OperationalError(f"Insertion failed with the following error: {e}") from e
    finally:
        connection.close()


def fetch_docs(database_path: str, ids: Iterable[int], table_name: str = "documents", column: str = "doc",
               connection: Connection = None) -> Dict[int, str]:
    """Batched ``SELECT {column} FROM {table} WHERE id IN (...)`` -> {id: text}."""
    ids = list(dict.fromkeys(int(i) for i in ids))
    if not ids:
        return {}
    own = connection is None
    con = connect_database(database_path) if own else connection
    try:
        out: Dict[int, str] = {}
        for i in range(0, len(ids), 900):  # sqlite's default variable limit is 999
            chunk = ids[i:i + 900]
            marks = ",".join("?" * len(chunk))
            for rid, text in con.execute(f"SELECT id, {column} FROM {table_name} WHERE id IN ({marks})", chunk):
                out[rid] = text
        return out
    finally:
        if own:
            con.close()
