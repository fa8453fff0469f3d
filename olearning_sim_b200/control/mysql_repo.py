"""MySQL backend of the persistent tables — what the reference deploys (SQLAlchemy-reflected MySQL tables:
ols_core/taskMgr/utils/utils.py:29-267 ``TaskTableRepo``, ols_core/resourceMgr/utils/utils.py ``ResTableRepo``,
ols_core/utils/repo_utils.py:19-345 ``SqlDataBase``).

``MySqlConnection`` adapts a DB-API connection (``pymysql`` when installed, otherwise the in-tree wire-protocol client
``utils/mysql_wire.py``) to the three things
``SqlDataBase`` needs from sqlite: ``execute(sql, params)`` with ``?`` placeholders and ``"identifier"`` quoting (rewritten
to ``%s`` / backticks), a cursor with ``rowcount`` / ``fetchone`` / ``fetchall`` / ``description``, and autocommit.  Like every
repo method of the reference (utils.py:85-97, repo_utils.py:89-106) a statement that fails with an operational error
re-creates the connection and is retried ONCE.  ``MySqlTaskTableRepo`` & co. are the table repos on that connection:

    repo = MySqlTaskTableRepo(host="db", user="ols", password=os.environ["OLS_DB_PW"], database="virtually_phone_DB")

No DSN or credential is ever hard-coded (the reference embeds them in source: task_runner.py:240, dispatcher.py:270).
"""
from __future__ import annotations

import re
import threading
from typing import Any, Callable, Dict, Optional, Sequence

from .repo import FLOW_COLUMNS, LOG_COLUMNS, RES_COLUMNS, TASK_COLUMNS, FlowTableRepo, LogTableRepo, ResTableRepo, SqlDataBase, TaskTableRepo

_IDENT = re.compile(r'"([A-Za-z_][A-Za-z0-9_]*)"')


def translate(sql: str) -> str:
    """sqlite dialect of control/repo.py -> MySQL: quoting, placeholders, auto-increment key, TEXT primary columns"""
    s = _IDENT.sub(r"`\1`", sql).replace("?", "%s")
    s = s.replace("_rowid INTEGER PRIMARY KEY AUTOINCREMENT", "_rowid BIGINT PRIMARY KEY AUTO_INCREMENT")
    s = re.sub(r"\bTEXT\b", "LONGTEXT", s)
    s = re.sub(r"\bREAL\b", "DOUBLE", s)
    return s


class MySqlConnection:
    def __init__(self, connect: Optional[Callable[[], Any]] = None, operational_errors: Sequence[type] = (), **kw: Any):
        if connect is None:
            try:
                import pymysql  # type: ignore
                operational_errors = (pymysql.err.OperationalError, pymysql.err.InterfaceError)
                connect = lambda: pymysql.connect(autocommit=True, **kw)       # noqa: E731
            except ImportError:
                # no pymysql: the in-tree wire-protocol client (utils/mysql_wire.py) provides the same DB-API subset
                from ..utils import mysql_wire
                operational_errors = (mysql_wire.OperationalError, mysql_wire.InterfaceError)
                connect = lambda: mysql_wire.connect(autocommit=True, **kw)    # noqa: E731
        self._connect, self._errors = connect, tuple(operational_errors) or (ConnectionError,)
        self._conn = connect()
        self.reconnects = 0

    def execute(self, sql: str, params: Sequence[Any] = ()):
        q = translate(sql)
        for attempt in (0, 1):
            try:
                cur = self._conn.cursor()
                cur.execute(q, tuple(params))
                if hasattr(self._conn, "commit"):
                    self._conn.commit()
                return cur
            except self._errors:
                if attempt:
                    raise
                try:
                    self._conn.close()
                except Exception:  # noqa: BLE001
                    pass
                self._conn = self._connect()                               # reconnect once, then retry the statement
                self.reconnects += 1


class _MySqlMixin:
    def _bind(self, table: str, columns, conn: MySqlConnection) -> None:
        self.path, self.table, self.columns = "mysql", table, list(columns)
        self._conn, self._lock, self._key = conn, threading.RLock(), f"mysql:{id(conn)}"
        self._create()


class MySqlDataBase(_MySqlMixin, SqlDataBase):
    def __init__(self, table: str, columns, conn: Optional[MySqlConnection] = None, **connect_kw: Any):
        self._bind(table, columns, conn or MySqlConnection(**connect_kw))


class MySqlTaskTableRepo(_MySqlMixin, TaskTableRepo):
    def __init__(self, conn: Optional[MySqlConnection] = None, **connect_kw: Any):
        self._bind("taskmgr_table", TASK_COLUMNS, conn or MySqlConnection(**connect_kw))


class MySqlResTableRepo(_MySqlMixin, ResTableRepo):
    def __init__(self, conn: Optional[MySqlConnection] = None, **connect_kw: Any):
        self._bind("resmgr_table", RES_COLUMNS, conn or MySqlConnection(**connect_kw))


class MySqlFlowTableRepo(_MySqlMixin, FlowTableRepo):
    def __init__(self, conn: Optional[MySqlConnection] = None, **connect_kw: Any):
        self._bind("deviceflow_table", FLOW_COLUMNS, conn or MySqlConnection(**connect_kw))


class MySqlLogTableRepo(_MySqlMixin, LogTableRepo):
    def __init__(self, conn: Optional[MySqlConnection] = None, **connect_kw: Any):
        self._bind("log_table", LOG_COLUMNS, conn or MySqlConnection(**connect_kw))
