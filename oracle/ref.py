"""TEST INFRASTRUCTURE — ctypes loader for oracle/kao_ref.c (the plain-C restatement of the
search path).  Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libkao_ref.so")


class RefProblem(C.Structure):
    _fields_ = [("P", C.c_int32), ("B", C.c_int32), ("R", C.c_int32), ("RF", C.c_int32),
                ("RFcur", C.c_int32),
                ("rack_of", C.c_void_p), ("wF", C.c_void_p), ("wL", C.c_void_p),
                ("rep_lo", C.c_void_p), ("rep_hi", C.c_void_p),
                ("ldr_lo", C.c_void_p), ("ldr_hi", C.c_void_p),
                ("rack_lo", C.c_void_p), ("rack_hi", C.c_void_p),
                ("ppr_lo", C.c_int32), ("ppr_hi", C.c_int32), ("cur", C.c_void_p)]


CFLAGS = "-O3 -march=native -fopenmp -fPIC -Wall -Wextra -std=c11"      # the stated CPU baseline is not a strawman build


def _host_stamp() -> str:
    """-march=native code is only valid on the CPU it was built on: the library is rebuilt when the host's
    instruction-set flags (or the build flags) differ from the ones it was built with."""
    flags = ""
    try:
        with open("/proc/cpuinfo") as f:
            for l in f:
                if l.startswith("flags"):
                    flags = " ".join(sorted(l.split(":", 1)[1].split()))
                    break
    except OSError:
        pass
    import hashlib

    return hashlib.sha256((CFLAGS + "|" + flags).encode()).hexdigest()


def build_flags() -> str:
    return "gcc " + CFLAGS


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "kao_ref.c")
    stamp_path = _SO + ".stamp"
    stamp = _host_stamp()
    try:
        with open(stamp_path) as f:
            same_host = f.read().strip() == stamp
    except OSError:
        same_host = False
    if force or not same_host or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B", "CFLAGS=" + CFLAGS])
        with open(stamp_path, "w") as f:
            f.write(stamp)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.kao_ref_pack.restype = C.c_uint64
        _lib.kao_ref_pack.argtypes = [C.c_int64, C.c_int64, C.c_uint32, C.c_int]
        _lib.kao_ref_search.restype = C.c_uint64
    return _lib


def _arr(a, dt):
    return np.ascontiguousarray(a, dtype=dt)


class Ref:
    """Holds the numpy buffers alive and exposes the C restatement on one oracle.model.Problem."""

    def __init__(self, pb):
        self.pb = pb
        self._keep = dict(
            rack_of=_arr(pb.rack_of, np.uint8), wF=_arr(pb.wF, np.uint16), wL=_arr(pb.wL, np.uint16),
            rep_lo=_arr(pb.rep_lo, np.int32), rep_hi=_arr(pb.rep_hi, np.int32),
            ldr_lo=_arr(pb.ldr_lo, np.int32), ldr_hi=_arr(pb.ldr_hi, np.int32),
            rack_lo=_arr(pb.rack_lo, np.int32), rack_hi=_arr(pb.rack_hi, np.int32),
            cur=_arr(pb.cur, np.int32))
        k = self._keep
        self.c = RefProblem(pb.P, pb.B, pb.R, pb.RF, pb.cur.shape[1],
                            *(k[n].ctypes.data for n in ("rack_of", "wF", "wL", "rep_lo", "rep_hi",
                                                        "ldr_lo", "ldr_hi", "rack_lo", "rack_hi")),
                            int(pb.ppr_lo), int(pb.ppr_hi), k["cur"].ctypes.data)
        self.W = lib().kao_ref_words(C.byref(self.c))
        if self.W < 1:
            raise ValueError("unsupported topology (more than 256 rack-aligned broker slots)")
        self.obj_bits = int(lib().kao_ref_obj_bits(C.byref(self.c)))     # cost-field width of this problem's keys

    def _p(self):
        return C.byref(self.c)

    def new_candidate(self):
        return np.zeros((self.pb.P, self.W), np.uint32), np.zeros(self.pb.P, np.uint8)

    def init_base(self):
        bits, ld = self.new_candidate()
        lib().kao_ref_init_base(self._p(), C.c_void_p(bits.ctypes.data), C.c_void_p(ld.ctypes.data))
        return bits, ld

    def evaluate(self, bits, ld):
        v, o = C.c_int64(), C.c_int64()
        lib().kao_ref_eval(self._p(), C.c_void_p(bits.ctypes.data), C.c_void_p(ld.ctypes.data),
                           C.byref(v), C.byref(o))
        return v.value, o.value

    def gen(self, bits, ld, seed, rnd, idx, round_size):
        ob, ol = self.new_candidate()
        lib().kao_ref_gen(self._p(), C.c_void_p(bits.ctypes.data), C.c_void_p(ld.ctypes.data),
                          C.c_uint64(seed), C.c_uint32(rnd), C.c_uint32(idx), C.c_uint32(round_size),
                          C.c_void_p(ob.ctypes.data), C.c_void_p(ol.ctypes.data))
        return ob, ol

    def candidate_keys(self, bits, ld, seed, rnd, round_size, idx_begin, count, nthreads=0):
        out = np.empty(count, np.uint64)
        lib().kao_ref_candidate_keys(self._p(), C.c_void_p(bits.ctypes.data),
                                     C.c_void_p(ld.ctypes.data), C.c_uint64(seed), C.c_uint32(rnd),
                                     C.c_uint32(round_size), C.c_uint32(idx_begin),
                                     C.c_uint32(count), C.c_void_p(out.ctypes.data),
                                     C.c_int(nthreads))
        return out

    def search(self, bits, ld, seed, first_round, rounds, round_size, nthreads=0):
        """In-place on (bits, ld).  Returns (last key, per-round keys)."""
        keys = np.zeros(rounds, np.uint64)
        last = lib().kao_ref_search(self._p(), C.c_void_p(bits.ctypes.data),
                                    C.c_void_p(ld.ctypes.data), C.c_uint64(seed),
                                    C.c_uint32(first_round), C.c_uint32(rounds),
                                    C.c_uint32(round_size), C.c_void_p(keys.ctypes.data),
                                    C.c_int(nthreads))
        return last, keys

    def decode(self, bits, ld):
        """bit-plane -> replica lists (dense broker indices, leader first)."""
        out = np.empty((self.pb.P, self.pb.RF), np.int32)
        lib().kao_ref_decode(self._p(), C.c_void_p(bits.ctypes.data), C.c_void_p(ld.ctypes.data),
                             C.c_void_p(out.ctypes.data))
        return out

    def encode(self, replicas):
        """replica lists (dense broker indices, leader first, -1 padded) -> bit-plane."""
        reps = np.ascontiguousarray(replicas, dtype=np.int32)
        bits, ld = self.new_candidate()
        lib().kao_ref_encode(self._p(), C.c_void_p(reps.ctypes.data), C.c_void_p(bits.ctypes.data),
                             C.c_void_p(ld.ctypes.data))
        return bits, ld

    @staticmethod
    def philox(ctr, key):
        c = (C.c_uint32 * 4)(*ctr); k = (C.c_uint32 * 2)(*key); o = (C.c_uint32 * 4)()
        lib().kao_ref_philox(c, k, o)
        return list(o)

    @staticmethod
    def max_threads():
        return lib().kao_ref_max_threads()


    def unpack_key(self, key):
        return unpack_key(key, self.obj_bits)

    def pack_key(self, viol, obj, idx):
        return int(lib().kao_ref_pack(int(viol), int(obj), int(idx), self.obj_bits))


def unpack_key(key, obj_bits):
    """packed key -> (violation, objective, index); obj_bits = Ref.obj_bits (docs/MODEL.md 3)."""
    key, omax = int(key), (1 << obj_bits) - 1
    return key >> (24 + obj_bits), omax - ((key >> 24) & omax), key & 0xFFFFFF
