"""ctypes access to the CPU oracle (oracle/_build/liboracle.so) -- tests only."""
import ctypes as C

import numpy as np

import util


class OSeqInfo(C.Structure):
    _fields_ = [("offset", C.c_uint64), ("size", C.c_uint64), ("hdr_offset", C.c_uint64),
                ("platform", C.c_int32), ("pad", C.c_int32)]


class OVolume(C.Structure):
    _fields_ = [("pac", C.POINTER(C.c_uint8)), ("db_size", C.c_uint64), ("nseq", C.c_uint64),
                ("seq", C.POINTER(OSeqInfo)), ("hdr", C.c_char_p), ("hdr_size", C.c_uint64)]


class OAlignment(C.Structure):
    _fields_ = [("ok", C.c_int), ("qoff", C.c_int), ("qend", C.c_int), ("toff", C.c_int), ("tend", C.c_int),
                ("ident_perc", C.c_double), ("qaln", C.POINTER(C.c_char)), ("taln", C.POINTER(C.c_char)),
                ("aln_size", C.c_int)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        util.ensure_oracle()
        L = C.CDLL(util.ORACLE_LIB)
        L.o_aligner_new.restype = C.c_void_p
        L.o_aligner_new.argtypes = [C.c_double]
        L.o_aligner_free.argtypes = [C.c_void_p]
        L.o_onc_align.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                  C.c_int, C.c_int, C.POINTER(OAlignment)]
        L.o_edlib_align.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_char_p, C.c_char_p,
                                    C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.o_edlib_align_dp.argtypes = [C.c_double, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_char_p, C.c_char_p,
                                       C.POINTER(C.c_int), C.POINTER(C.c_int)]
        _lib = L
    return _lib


def strings_to_ops(qaln, taln):
    """gapped strings -> op bytes of include/necat_b200.h (0 match, 1 query only, 2 target only, 3 mismatch)"""
    q = np.frombuffer(qaln, np.uint8)
    t = np.frombuffer(taln, np.uint8)
    ops = np.where(q == t, 0, 3).astype(np.uint8)
    ops[t == ord("-")] = 1
    ops[q == ord("-")] = 2
    return ops


def onc_align(aligner, query, qstart, target, tstart, min_align_size, tail, block=512):
    """query/target: uint8 code arrays.  Returns dict(ok, qoff, qend, toff, tend, ops)."""
    L = lib()
    q = np.ascontiguousarray(query, np.uint8)
    t = np.ascontiguousarray(target, np.uint8)
    A = OAlignment()
    L.o_onc_align(aligner, q.ctypes.data, qstart, q.size, t.ctypes.data, tstart, t.size, block, min_align_size, tail,
                  C.byref(A))
    n = A.aln_size
    qa = C.string_at(A.qaln, n)
    ta = C.string_at(A.taln, n)
    return dict(ok=A.ok, qoff=A.qoff, qend=A.qend, toff=A.toff, tend=A.tend, ident=A.ident_perc,
                ops=strings_to_ops(qa, ta))
