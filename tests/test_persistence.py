"""Host-side tuple formats of the reference's persistent index (persistence.py).  CPU only: byte
layouts, round trips, and the reload drivers against a recording stand-in for the index classes."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import importlib.util  # noqa: E402

_spec = importlib.util.spec_from_file_location("mmidx_persistence", os.path.join(ROOT, "multimedia-indexing_b200", "persistence.py"))
P = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(P)


def test_integer_binding_orders_like_signed_ints():
    assert P.int_key(0) == b"\x80\x00\x00\x00"
    assert P.int_key(1) == b"\x80\x00\x00\x01"
    assert P.int_key(-1) == b"\x7f\xff\xff\xff"
    assert P.int_key(2**31 - 1) == b"\xff\xff\xff\xff"
    vals = [-5, -1, 0, 1, 7, 300, 2**31 - 1]
    keys = [P.int_key(v) for v in vals]
    assert keys == sorted(keys)  # byte order == numeric order, what JE's btree relies on
    assert [P.key_int(k) for k in keys] == vals


def test_byte_codes_are_the_raw_unsigned_code_on_disk():
    # in memory the reference keeps (byte)(c - 128) (PQ.java:555); TupleOutput.writeByte stores b ^ 0x80
    code = np.array([0, 1, 127, 128, 200, 255], dtype=np.uint8)
    java_bytes = (code.astype(np.int16) - 128).astype(np.int8)
    on_disk = bytes(((int(b) ^ 0x80) & 0xFF) for b in java_bytes)
    assert P.encode_adc(code) == on_disk == code.tobytes()
    assert (P.decode_adc(on_disk, 6) == code).all()


def test_ivfadc_value_layout():
    v = P.encode_ivfadc(5, np.array([9, 255, 0, 128], dtype=np.uint8))
    assert v == b"\x80\x00\x00\x05" + bytes([9, 255, 0, 128])
    l, c = P.decode_ivfadc(v, 4)
    assert l == 5 and (c == [9, 255, 0, 128]).all()
    # ks > 256: big-endian shorts with the sign bit flipped, no -128 offset (IVFPQ.java:440-443)
    v = P.encode_ivfadc(8191, np.array([0, 1, 999, 65535 >> 1], dtype=np.uint16), ks=1000)
    assert v[:4] == b"\x80\x00\x1f\xff" and v[4:8] == b"\x80\x00\x80\x01"
    l, c = P.decode_ivfadc(v, 4, ks=1000)
    assert l == 8191 and (c == [0, 1, 999, 32767]).all()


def test_vlad_value_is_big_endian_raw_doubles():
    x = np.array([1.0, -2.5, 0.0, 1e-300])
    b = P.encode_vlad(x)
    assert b[:8] == b"\x3f\xf0\x00\x00\x00\x00\x00\x00" and len(b) == 32
    assert (P.decode_vlad(b, 4) == x).all()
    assert P.entry_string(P.string_entry("img_éñ.jpg")) == "img_éñ.jpg" and P.string_entry("a") == b"a\x00"


class _Recorder:
    """records what the reload drivers hand to the index classes"""

    def __init__(self, m=4, ks=256, d=3, loaded=0):
        self.numSubVectors, self.numProductCentroids, self.vectorLength = m, ks, d
        self.calls, self._n = [], loaded

    def getLoadCounter(self):
        return self._n

    def indexPQCodes(self, ids, *a):
        self.calls.append(tuple(np.array(x) for x in a))
        self._n += len(a[-1])

    def indexVectors(self, ids, X):
        self.calls.append((np.array(X),))
        self._n += len(X)


def test_reload_drivers_round_trip_and_check_key_order():
    rng = np.random.default_rng(0)
    lists = rng.integers(0, 50, size=1000).astype(np.int32)
    codes = rng.integers(0, 256, size=(1000, 4)).astype(np.uint8)
    recs = list(P.ivfpq_records(lists, codes))
    assert recs[3][0] == P.int_key(3) and len(recs[0][1]) == 8
    ix = _Recorder()
    assert P.load_ivfpq(ix, recs, batch=300) == 1000 and len(ix.calls) == 4
    assert (np.concatenate([c[0] for c in ix.calls]) == lists).all()
    assert (np.concatenate([c[1] for c in ix.calls]) == codes).all()
    with pytest.raises(ValueError):  # a gap in the keys would silently shift every later iid in the reference
        P.load_ivfpq(_Recorder(), recs[:10] + recs[11:])
    # appending to a partly loaded index: keys continue at loadCounter
    more = list(P.ivfpq_records(lists[:5], codes[:5], first_iid=1000))
    assert P.load_ivfpq(ix, more) == 5
    pq = _Recorder()
    assert P.load_pq(pq, [(P.int_key(i), P.encode_adc(codes[i])) for i in range(100)], batch=64) == 100
    assert (np.concatenate([c[0] for c in pq.calls]) == codes[:100]).all()
    X = rng.normal(size=(20, 3))
    lin = _Recorder()
    assert P.load_linear(lin, [(P.int_key(i), P.encode_vlad(X[i])) for i in range(20)], batch=8) == 20
    assert (np.concatenate([c[0] for c in lin.calls]) == X).all()
