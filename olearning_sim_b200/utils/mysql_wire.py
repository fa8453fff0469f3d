"""Minimal MySQL client on the wire protocol (text protocol, ``mysql_native_password`` and ``caching_sha2_password``) — the DB-API subset ``control/mysql_repo.py`` needs, without ``pymysql``.

The reference keeps its task / resource / flow / log tables in MySQL through SQLAlchemy + pymysql
(ols_core/utils/repo_utils.py:19-345, taskMgr/utils/utils.py:29-267).  Neither is installable here; ``MySqlConnection``
falls back to this module:

    conn = mysql_wire.connect(host=..., port=3306, user=..., password=..., database=..., autocommit=True)
    cur = conn.cursor(); cur.execute("SELECT a FROM t WHERE b = %s", ("x",)); cur.fetchall(); cur.rowcount; cur.description

Protocol (MySQL 5.7 / 8.0 "Client/Server Protocol"): 4-byte packet header (3-byte little-endian length + sequence id),
HandshakeV10 -> HandshakeResponse41 (+ AuthSwitchRequest), COM_QUERY with OK / ERR / text result sets (with or without
CLIENT_DEPRECATE_EOF), COM_PING, COM_QUIT.  Parameters are interpolated client side with MySQL string escaping, as pymysql
does.  ``caching_sha2_password`` full authentication (the password is not in the server's cache yet) goes through the RSA
public-key exchange (``cryptography``).  Not implemented: TLS, compression, prepared statements.

tests/test_round2_parity.py runs the table repos over TCP against an in-process server that speaks this protocol.
"""
from __future__ import annotations

import hashlib
import socket
import struct
from typing import Any, List, Optional, Sequence, Tuple

CLIENT_LONG_PASSWORD, CLIENT_LONG_FLAG, CLIENT_CONNECT_WITH_DB = 0x1, 0x4, 0x8
CLIENT_PROTOCOL_41, CLIENT_TRANSACTIONS, CLIENT_SECURE_CONNECTION = 0x200, 0x2000, 0x8000
CLIENT_MULTI_RESULTS, CLIENT_PLUGIN_AUTH, CLIENT_DEPRECATE_EOF = 0x20000, 0x80000, 0x1000000
UTF8MB4 = 45


class Error(Exception):
    pass


class InterfaceError(Error):
    """socket / framing problems: the connection is unusable"""


class OperationalError(Error):
    """server-side operational errors (lost connection, access denied, unknown database ...)"""


class ProgrammingError(Error):
    """SQL errors (syntax, unknown table / column, duplicate key ...)"""


def scramble_native(password: str, salt: bytes) -> bytes:
    """mysql_native_password: SHA1(pw) XOR SHA1(salt + SHA1(SHA1(pw)))"""
    if not password:
        return b""
    p1 = hashlib.sha1(password.encode()).digest()
    p2 = hashlib.sha1(p1).digest()
    p3 = hashlib.sha1(salt + p2).digest()
    return bytes(a ^ b for a, b in zip(p1, p3))


def scramble_sha2(password: str, salt: bytes) -> bytes:
    """caching_sha2_password fast path: SHA256(pw) XOR SHA256(SHA256(SHA256(pw)) + salt)"""
    if not password:
        return b""
    p1 = hashlib.sha256(password.encode()).digest()
    p2 = hashlib.sha256(p1).digest()
    p3 = hashlib.sha256(p2 + salt).digest()
    return bytes(a ^ b for a, b in zip(p1, p3))


def rsa_encrypt_password(password: str, salt: bytes, pem: bytes) -> bytes:
    """caching_sha2_password / sha256_password full authentication over an unencrypted connection: (password + NUL) XOR
    salt (repeated), encrypted with the server's RSA public key, OAEP / SHA-1 padding (what libmysqlclient does)."""
    try:
        from cryptography.hazmat.primitives import hashes, serialization
        from cryptography.hazmat.primitives.asymmetric import padding
    except ImportError as e:  # pragma: no cover - the package is part of this image
        raise OperationalError("(2061) caching_sha2_password full authentication needs the 'cryptography' package") from e
    pw = password.encode() + b"\0"
    x = bytes(b ^ salt[i % len(salt)] for i, b in enumerate(pw))
    pub = serialization.load_pem_public_key(pem)
    return pub.encrypt(x, padding.OAEP(mgf=padding.MGF1(algorithm=hashes.SHA1()), algorithm=hashes.SHA1(), label=None))


def lenenc_int(buf: bytes, pos: int) -> Tuple[Optional[int], int]:
    c = buf[pos]
    if c < 0xFB:
        return c, pos + 1
    if c == 0xFB:
        return None, pos + 1
    if c == 0xFC:
        return struct.unpack_from("<H", buf, pos + 1)[0], pos + 3
    if c == 0xFD:
        return int.from_bytes(buf[pos + 1:pos + 4], "little"), pos + 4
    return struct.unpack_from("<Q", buf, pos + 1)[0], pos + 9


def put_lenenc(n: int) -> bytes:
    if n < 0xFB:
        return bytes([n])
    if n < 1 << 16:
        return b"\xfc" + struct.pack("<H", n)
    if n < 1 << 24:
        return b"\xfd" + n.to_bytes(3, "little")
    return b"\xfe" + struct.pack("<Q", n)


_ESC = {0: "\\0", 8: "\\b", 9: "\\t", 10: "\\n", 13: "\\r", 26: "\\Z", 34: '\\"', 39: "\\'", 92: "\\\\"}


def escape(v: Any) -> str:
    """literal of a Python value in MySQL syntax"""
    if v is None:
        return "NULL"
    if isinstance(v, bool):
        return "1" if v else "0"
    if isinstance(v, (int, float)):
        return repr(v)
    if isinstance(v, (bytes, bytearray)):
        return "x'" + bytes(v).hex() + "'"
    return "'" + str(v).translate(_ESC) + "'"


def interpolate(sql: str, params: Sequence[Any]) -> str:
    if not params:
        return sql
    parts = sql.split("%s")
    if len(parts) - 1 != len(params):
        raise ProgrammingError(f"statement has {len(parts) - 1} placeholders, {len(params)} parameters given")
    out = [parts[0]]
    for p, rest in zip(params, parts[1:]):
        out += [escape(p), rest]
    return "".join(out)


# column types whose text value converts to int / float (everything else stays str; BLOBs bytes)
_INT_TYPES, _FLOAT_TYPES, _BLOB_TYPES = {1, 2, 3, 8, 9, 13}, {0, 4, 5, 246}, {249, 250, 251, 252}


class Cursor:
    def __init__(self, conn: "Connection"):
        self._c = conn
        self.rowcount, self.lastrowid, self.description = -1, None, None
        self._rows: List[tuple] = []
        self._i = 0

    def execute(self, sql: str, params: Sequence[Any] = ()) -> "Cursor":
        self.rowcount, self.lastrowid, self.description, self._rows, self._i = self._c._query(interpolate(sql, params))
        return self

    def fetchone(self) -> Optional[tuple]:
        if self._i >= len(self._rows):
            return None
        self._i += 1
        return self._rows[self._i - 1]

    def fetchall(self) -> List[tuple]:
        out, self._i = self._rows[self._i:], len(self._rows)
        return out

    def close(self) -> None:
        self._rows = []


class Connection:
    def __init__(self, host: str = "127.0.0.1", port: int = 3306, user: str = "", password: str = "", database: str = "",
                 autocommit: bool = True, connect_timeout: float = 10.0, **_ignored: Any):
        self._seq = 0
        try:
            self._s = socket.create_connection((host, int(port)), timeout=connect_timeout)
        except OSError as e:
            raise OperationalError(f"(2003) can't connect to MySQL server on {host}:{port} ({e})") from None
        self._s.settimeout(None)
        self._caps = 0
        try:
            self._handshake(user, password, database)
            if autocommit:
                self._query("SET autocommit=1")
        except Exception:
            self._s.close()
            raise

    # ---- framing -------------------------------------------------------------------------------------------------
    def _recv_exact(self, n: int) -> bytes:
        buf = b""
        while len(buf) < n:
            try:
                chunk = self._s.recv(n - len(buf))
            except OSError as e:
                raise OperationalError(f"(2013) lost connection to MySQL server during query ({e})") from None
            if not chunk:
                raise OperationalError("(2013) lost connection to MySQL server during query")
            buf += chunk
        return buf

    def _read(self) -> bytes:
        out = b""
        while True:
            hdr = self._recv_exact(4)
            n, seq = int.from_bytes(hdr[:3], "little"), hdr[3]
            self._seq = (seq + 1) & 0xFF
            out += self._recv_exact(n)
            if n < 0xFFFFFF:
                return out

    def _write(self, payload: bytes) -> None:
        try:
            while True:
                chunk, payload = payload[:0xFFFFFF], payload[0xFFFFFF:]
                self._s.sendall(len(chunk).to_bytes(3, "little") + bytes([self._seq]) + chunk)
                self._seq = (self._seq + 1) & 0xFF
                if len(chunk) < 0xFFFFFF:
                    return
        except OSError as e:
            raise OperationalError(f"(2006) MySQL server has gone away ({e})") from None

    @staticmethod
    def _raise(pkt: bytes) -> None:
        code = struct.unpack_from("<H", pkt, 1)[0]
        msg = pkt[9:].decode("utf-8", "replace") if len(pkt) > 3 and pkt[3:4] == b"#" else pkt[3:].decode("utf-8", "replace")
        cls = OperationalError if code in (1040, 1044, 1045, 1049, 1053, 1152, 1205, 1213, 2006, 2013) else ProgrammingError
        raise cls(f"({code}) {msg}")

    # ---- connection phase ------------------------------------------------------------------------------------------
    def _handshake(self, user: str, password: str, database: str) -> None:
        pkt = self._read()
        if pkt[0] == 0xFF:
            self._raise(pkt)
        if pkt[0] != 10:
            raise InterfaceError(f"unsupported handshake protocol {pkt[0]}")
        pos = pkt.index(b"\0", 1) + 1                      # server version
        pos += 4                                           # thread id
        salt = pkt[pos:pos + 8]
        pos += 9
        caps = struct.unpack_from("<H", pkt, pos)[0]
        pos += 2
        plugin = "mysql_native_password"
        if len(pkt) > pos:
            pos += 3                                       # charset, status
            caps |= struct.unpack_from("<H", pkt, pos)[0] << 16
            pos += 2
            alen = pkt[pos]
            pos += 11                                      # auth data length + 10 reserved
            n2 = max(13, alen - 8)
            salt += pkt[pos:pos + n2].rstrip(b"\0")
            pos += n2
            if caps & CLIENT_PLUGIN_AUTH and pos < len(pkt):
                end = pkt.find(b"\0", pos)
                plugin = pkt[pos:end if end >= 0 else len(pkt)].decode()
        mine = (CLIENT_LONG_PASSWORD | CLIENT_LONG_FLAG | CLIENT_PROTOCOL_41 | CLIENT_TRANSACTIONS | CLIENT_SECURE_CONNECTION |
                CLIENT_MULTI_RESULTS | CLIENT_PLUGIN_AUTH | (CLIENT_CONNECT_WITH_DB if database else 0))
        self._caps = mine & (caps | CLIENT_CONNECT_WITH_DB)          # no DEPRECATE_EOF requested: classic EOF framing
        auth = scramble_sha2(password, salt) if plugin == "caching_sha2_password" else scramble_native(password, salt)
        resp = struct.pack("<IIB23x", self._caps, 1 << 24, UTF8MB4) + user.encode() + b"\0" + bytes([len(auth)]) + auth
        if database:
            resp += database.encode() + b"\0"
        resp += plugin.encode() + b"\0"
        self._write(resp)
        while True:
            pkt = self._read()
            if pkt[0] == 0x00:
                return
            if pkt[0] == 0xFF:
                self._raise(pkt)
            if pkt[0] == 0xFE:                              # AuthSwitchRequest: plugin name + new salt
                end = pkt.index(b"\0", 1)
                plugin, salt = pkt[1:end].decode(), pkt[end + 1:].rstrip(b"\0")
                if plugin == "mysql_native_password":
                    self._write(scramble_native(password, salt))
                elif plugin == "caching_sha2_password":
                    self._write(scramble_sha2(password, salt))
                else:
                    raise OperationalError(f"(2059) authentication plugin '{plugin}' is not supported")
                continue
            if pkt[0] == 0x01:                              # AuthMoreData
                if pkt[1:2] == b"\x03":                     # caching_sha2_password: fast authentication succeeded
                    continue
                if pkt[1:2] == b"\x04":                     # full authentication: password under the server's RSA key
                    self._write(b"\x02")                    # request the public key
                    key = self._read()
                    if key[0] != 0x01:
                        raise InterfaceError("expected the server's RSA public key")
                    self._write(rsa_encrypt_password(password, salt, key[1:]))
                    continue
            raise InterfaceError(f"unexpected packet 0x{pkt[0]:02x} during authentication")

    # ---- command phase ---------------------------------------------------------------------------------------------
    def _query(self, sql: str):
        self._seq = 0
        self._write(b"\x03" + sql.encode("utf-8"))
        pkt = self._read()
        if pkt[0] == 0xFF:
            self._raise(pkt)
        if pkt[0] == 0x00:                                  # OK: affected rows, last insert id
            aff, pos = lenenc_int(pkt, 1)
            last, pos = lenenc_int(pkt, pos)
            return aff, last, None, [], 0
        ncol, _ = lenenc_int(pkt, 0)
        desc, types = [], []
        for _ in range(ncol or 0):
            col = self._read()
            pos, fields = 0, []
            for _k in range(6):                             # catalog, schema, table, org_table, name, org_name
                n, pos = lenenc_int(col, pos)
                fields.append(col[pos:pos + (n or 0)])
                pos += n or 0
            pos += 1                                        # length of the fixed fields (0x0c)
            charset, length, ctype = struct.unpack_from("<HIB", col, pos)
            desc.append((fields[4].decode(), ctype, None, length, None, None, None))
            types.append((ctype, charset))
        if not self._caps & CLIENT_DEPRECATE_EOF:
            eof = self._read()
            if eof[0] != 0xFE:
                raise InterfaceError("expected EOF after the column definitions")
        rows: List[tuple] = []
        while True:
            pkt = self._read()
            if pkt[0] == 0xFF:
                self._raise(pkt)
            if pkt[0] == 0xFE and len(pkt) < 9:
                break
            pos, row = 0, []
            for ctype, charset in types:
                n, pos = lenenc_int(pkt, pos)
                if n is None:
                    row.append(None)
                    continue
                raw = pkt[pos:pos + n]
                pos += n
                if ctype in _INT_TYPES:
                    row.append(int(raw))
                elif ctype in _FLOAT_TYPES:
                    row.append(float(raw))
                elif ctype in _BLOB_TYPES and charset == 63:
                    row.append(bytes(raw))
                else:
                    row.append(raw.decode("utf-8", "replace"))
            rows.append(tuple(row))
        return len(rows), None, desc, rows, 0

    # ---- DB-API surface --------------------------------------------------------------------------------------------
    def cursor(self) -> Cursor:
        return Cursor(self)

    def commit(self) -> None:
        self._query("COMMIT")

    def rollback(self) -> None:
        self._query("ROLLBACK")

    def ping(self) -> bool:
        self._seq = 0
        self._write(b"\x0e")
        return self._read()[0] == 0x00

    def close(self) -> None:
        try:
            self._seq = 0
            self._write(b"\x01")
        except Error:
            pass
        try:
            self._s.close()
        except OSError:
            pass


def connect(**kw: Any) -> Connection:
    return Connection(**kw)
