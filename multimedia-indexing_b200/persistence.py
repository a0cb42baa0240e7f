"""Host-side tuple formats of the reference's persistent index, so a host that
keeps the reference's Berkeley DB JE databases can persist what `mmidx_add` returns and rebuild the GPU lists with
`mmidx_add_codes` exactly as `IVFPQ.loadIndexInMemory` (IVFPQ.java:680-728) / `PQ.loadIndexInMemory` (PQ.java:436-521) /
`Linear.loadIndexInMemory` (Linear.java:191-242) do from a cursor scan.

Only the byte layouts of keys and values are restated here (plain Python, no GPU work); the JE log files themselves stay
a JE concern.  The encodings are those of JE's `TupleOutput` / `IntegerBinding` / `StringBinding` as recalled in
SURVEY.md A.5 -- the jars are not vendored in the reference, so they are UNVERIFIED against a real index:

  key    IntegerBinding          4 bytes big-endian, sign bit flipped             (ASS.java:576-587)
  "ivfadc" value                 writeInt(listId) = (v ^ 0x80000000) big-endian, then the code   (IVFPQ.java:760-772)
  "adc" value                    the code                                          (PQ.java:493-496)
      code, ks <= 256            m x writeByte(b) = b ^ 0x80 with the in-memory byte b = c - 128 (PQ.java:555): the raw
                                 unsigned code c on disk
      code, ks  > 256            m x writeShort(s) = (s ^ 0x8000) big-endian with s = c (no offset, IVFPQ.java:440-443)
  "vlad" value                   d x writeDouble = raw IEEE bits, big-endian       (Linear.java:233-236)
  ids                            StringBinding: UTF-8 bytes + NUL                  (ASS.java:576-587)
"""
import struct

import numpy as np


def int_key(iid):
    """IntegerBinding.intToEntry"""
    return struct.pack(">I", (int(iid) ^ 0x80000000) & 0xFFFFFFFF)


def key_int(b):
    (u,) = struct.unpack(">I", bytes(b[:4]))
    v = u ^ 0x80000000
    return v - (1 << 32) if v & 0x80000000 else v


def string_entry(s):
    """StringBinding.stringToEntry"""
    return s.encode("utf-8") + b"\x00"


def entry_string(b):
    b = bytes(b)
    return b[: b.index(b"\x00")].decode("utf-8") if b"\x00" in b else b.decode("utf-8")


def _encode_code(code, ks):
    code = np.asarray(code)
    if ks <= 256:
        return np.ascontiguousarray(code, dtype=np.uint8).tobytes()  # ((c - 128) ^ 0x80) & 0xff == c
    return (np.asarray(code, dtype=np.uint16) ^ np.uint16(0x8000)).astype(">u2").tobytes()


def _decode_code(buf, m, ks):
    if ks <= 256:
        return np.frombuffer(buf, dtype=np.uint8, count=m).copy()
    return (np.frombuffer(buf, dtype=">u2", count=m).astype(np.uint16) ^ np.uint16(0x8000)).astype(np.uint16)


def encode_ivfadc(list_id, code, ks=256):
    """value of the "ivfadc" database for one vector: (listId, raw codes 0..ks-1)"""
    return struct.pack(">I", (int(list_id) ^ 0x80000000) & 0xFFFFFFFF) + _encode_code(code, ks)


def decode_ivfadc(value, m, ks=256):
    value = bytes(value)
    return key_int(value[:4]), _decode_code(value[4:], m, ks)


def encode_adc(code, ks=256):
    """value of the "adc" database (flat PQ index)"""
    return _encode_code(code, ks)


def decode_adc(value, m, ks=256):
    return _decode_code(bytes(value), m, ks)


def encode_vlad(vector):
    """value of the "vlad" database (Linear index): d big-endian doubles"""
    return np.ascontiguousarray(vector, dtype=np.float64).astype(">f8").tobytes()


def decode_vlad(value, d):
    return np.frombuffer(bytes(value), dtype=">f8", count=d).astype(np.float64)


def ivfpq_records(lists, codes, ks=256, first_iid=0):
    """(key, value) pairs in iid order for what indexVectors(..., return_codes=True) / mmidx_add returned -- the tuples
    IVFPQ.appendPersistentIndex (IVFPQ.java:760-792) writes."""
    for i, (l, c) in enumerate(zip(lists, codes)):
        yield int_key(first_iid + i), encode_ivfadc(l, c, ks)


def load_ivfpq(index, records, batch=1 << 16):
    """Rebuilds an IVFPQ index from "ivfadc" tuples in key order (IVFPQ.loadIndexInMemory IVFPQ.java:680-728): the iid
    is the position of the record, exactly as the reference's cursor scan assigns `counter`.  Returns the number loaded."""
    m, ks = index.numSubVectors, index.numProductCentroids
    ls, cs, n, expect = [], [], 0, index.getLoadCounter()

    def flush():
        if ls:
            index.indexPQCodes(None, np.asarray(ls, dtype=np.int32), np.stack(cs))
            ls.clear()
            cs.clear()

    for key, value in records:
        if key_int(key) != expect + n:
            raise ValueError(f"record {n}: key {key_int(key)} breaks the dense iid order the index relies on")
        l, c = decode_ivfadc(value, m, ks)
        ls.append(l)
        cs.append(c)
        n += 1
        if len(ls) >= batch:
            flush()
    flush()
    return n


def load_pq(index, records, batch=1 << 16):
    """PQ.loadIndexInMemory (PQ.java:436-521) from "adc" tuples in key order."""
    m, ks = index.numSubVectors, index.numProductCentroids
    cs, n, expect = [], 0, index.getLoadCounter()
    for key, value in records:
        if key_int(key) != expect + n:
            raise ValueError(f"record {n}: key {key_int(key)} breaks the dense iid order the index relies on")
        cs.append(decode_adc(value, m, ks))
        n += 1
        if len(cs) >= batch:
            index.indexPQCodes(None, np.stack(cs))
            cs.clear()
    if cs:
        index.indexPQCodes(None, np.stack(cs))
    return n


def load_linear(index, records, batch=1 << 14):
    """Linear.loadIndexInMemory (Linear.java:191-242) from "vlad" tuples in key order."""
    d = index.vectorLength
    vs, n, expect = [], 0, index.getLoadCounter()
    for key, value in records:
        if key_int(key) != expect + n:
            raise ValueError(f"record {n}: key {key_int(key)} breaks the dense iid order the index relies on")
        vs.append(decode_vlad(value, d))
        n += 1
        if len(vs) >= batch:
            index.indexVectors(None, np.stack(vs))
            vs.clear()
    if vs:
        index.indexVectors(None, np.stack(vs))
    return n
