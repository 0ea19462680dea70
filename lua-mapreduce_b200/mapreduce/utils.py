"""Constants and the text wire format of mapreduce/utils.lua (shuffle half only)."""
import re

# mapreduce/utils.lua:33-56
class STATUS:
    WAITING, RUNNING, BROKEN, FINISHED, WRITTEN, FAILED = range(6)


class TASK_STATUS:
    WAIT, MAP, REDUCE, FINISHED = "WAIT", "MAP", "REDUCE", "FINISHED"


MAX_WORKER_RETRIES = 3
MAX_JOB_RETRIES = 3
MAX_MAP_RESULT = 5000
MAX_TASKFN_VALUE_SIZE = 16 * 1024
STORAGES = ("gridfs", "shared", "sshfs", "hbm")


def get_storage_from(s, new=False):
    """mapreduce/utils.lua:273-285: "<kind>[:/abs/path]" -> (kind, path).  "hbm" is the new
    kind; the Mongo/NFS/scp kinds of the reference are not provided by this package."""
    s = s or "hbm"
    m = re.match(r"^([^:]+):(/.*)$", s)
    storage, path = (m.group(1), m.group(2)) if m else (s, "/hbm")
    if storage not in STORAGES:
        raise ValueError("Given incorrect storage %s" % storage)  # fs.lua:205-206
    if storage != "hbm":
        raise ValueError("Given incorrect storage %s: this build only provides 'hbm'" % storage)
    return storage, path


def escape(v):
    """mapreduce/utils.lua:100-112: numbers via "%.14g", strings via Lua 5.2 %q with the
    backslash-newline pair rewritten to \\n.  Returns bytes."""
    if isinstance(v, bool):
        return b"true" if v else b"false"
    if isinstance(v, (int, float)):
        return (b"%.14g" % v)
    out = bytearray(b'"')
    b = v.encode() if isinstance(v, str) else bytes(v)
    for i, c in enumerate(b):
        if c in (0x22, 0x5C):
            out += bytes((0x5C, c))
        elif c == 0x0A:
            out += b"\\n"
        elif c == 0 or c < 32 or c == 127:
            nxt = b[i + 1] if i + 1 < len(b) else 0
            out += (b"\\%03d" if 48 <= nxt <= 57 else b"\\%d") % c
        else:
            out.append(c)
    out += b'"'
    return bytes(out)


def serialize_table_ipairs(vals):
    """mapreduce/utils.lua:114-120"""
    return b"{" + b",".join(escape(v) for v in vals) + b"}"


def result_line(key, values):
    """job.lua:272-273: one line of a result.P<kk> file (interchange/export format)."""
    return b"return " + escape(key) + b"," + serialize_table_ipairs(values) + b"\n"


def count_digits(n):
    """mapreduce/server.lua:134-145"""
    assert n >= 0, "Only valid for positive integers"
    return max(1, len(str(int(n))))
