"""Test infrastructure (oracle/): the Python half of the OpenCV bridge of oracle/ref_stubs_detect/cv_detect.h.  The compiled reference (oracle/_ref/libref_detect.so)
forwards every cv:: call of its detection path to the callback made here, which runs the REAL OpenCV function (cv2) on the same bytes and hands the
result back through ref_cv_out.  Constants arrive with OpenCV's own numeric values.  An exception inside a ctypes callback cannot propagate: it is
recorded in `errors`, which every test asserts to be empty."""
import ctypes as C

import numpy as np


class BridgeMat(C.Structure):
    _fields_ = [("rows", C.c_int32), ("cols", C.c_int32), ("type", C.c_int32), ("channels", C.c_int32), ("step", C.c_int64), ("data", C.c_void_p)]


BRIDGE = C.CFUNCTYPE(None, C.c_char_p, C.POINTER(BridgeMat), C.POINTER(BridgeMat), C.c_void_p, C.POINTER(C.c_double), C.c_int32)


def _array(m):
    """A copy of an 8-bit or float32 cv::Mat (possibly a view with a larger row step) as (rows, cols) or (rows, cols, channels)."""
    m = m.contents
    depth = m.type & 7
    if depth not in (0, 5):
        raise TypeError(f"bridge: only 8-bit and float32 matrices cross (type {m.type})")
    dt, es = (np.uint8, 1) if depth == 0 else (np.float32, 4)
    row = m.cols * m.channels * es
    if m.rows == 0 or row == 0:
        return np.zeros((m.rows, m.cols) if m.channels == 1 else (m.rows, m.cols, m.channels), dt)
    size = (m.rows - 1) * m.step + row
    flat = np.frombuffer((C.c_ubyte * size).from_address(m.data), np.uint8)
    a = np.lib.stride_tricks.as_strided(flat, (m.rows, row), (m.step, 1)).copy().view(dt)
    return a if m.channels == 1 else a.reshape(m.rows, m.cols, m.channels)


def _convert_8u(cv2, a):
    """cv::Mat::convertTo(dst, CV_8U) of a float32 matrix with OpenCV's own rounding and saturation: cv2.add with a zero matrix and dtype CV_8U runs the same
    saturate_cast<uchar>(float) (cvRound: round half to even, clamp to 0 ... 255) -- cv2 has no direct convertTo binding."""
    return cv2.add(a, np.zeros_like(a), dtype=cv2.CV_8U)


class Bridge:
    def __init__(self, lib):
        import cv2
        self.cv2 = cv2
        self.lib = lib
        self.errors = []
        self.calls = []
        lib.ref_cv_out.restype = C.c_void_p
        lib.ref_cv_out.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
        self._cb = BRIDGE(self._call)                      # kept alive with the object
        lib.ref_cv_set_bridge(self._cb)

    def _call(self, op, a, b, out, params, n):
        try:
            cv2 = self.cv2
            op = op.decode()
            p = [params[i] for i in range(n)]
            A = _array(a) if a else None
            B = _array(b) if b else None
            self.calls.append(op)
            if op == "cvtColor":
                r = cv2.cvtColor(A, int(p[0]))
            elif op == "absdiff":
                r = cv2.absdiff(A, B)
            elif op == "subtract":
                r = cv2.subtract(A, B)
            elif op == "subtract_scalar_mat":
                if A.ndim != 2:
                    raise ValueError("scalar - matrix on a multi-channel matrix")
                r = cv2.subtract(p[0], A)                                  # cv::subtract(Scalar(255), src) on a single-channel matrix
            elif op == "threshold":
                r = cv2.threshold(A, p[0], p[1], int(p[2]))[1]
            elif op == "blur":
                r = cv2.blur(A, (int(p[0]), int(p[1])))
            elif op == "inRange":
                r = cv2.inRange(A, p[0], p[1])
            elif op == "adaptiveThreshold":
                r = cv2.adaptiveThreshold(A, p[0], int(p[1]), int(p[2]), int(p[3]), p[4])
            elif op == "dilate":
                r = cv2.dilate(A, B)
            elif op == "erode":
                r = cv2.erode(A, B)
            elif op == "bitwise_and":
                r = cv2.bitwise_and(A, B)
            elif op == "bitwise_or":
                r = cv2.bitwise_or(A, B)
            elif op == "add":
                r = cv2.add(A, B)
            elif op == "max":
                r = cv2.max(A, B)
            elif op == "min":
                r = cv2.min(A, B)
            elif op == "divide_mat_scalar":
                r = cv2.divide(A, (p[0], p[0], p[0], p[0]))                 # cv::divide(src, double): the 1 x 1 scalar divides every channel
            elif op == "convertTo_8u":
                r = _convert_8u(cv2, A)
            elif op == "equalizeHist":
                r = cv2.equalizeHist(A)
            elif op == "getStructuringElement":
                r = cv2.getStructuringElement(int(p[0]), (int(p[1]), int(p[2])), (int(p[3]), int(p[4])))
            else:
                raise NotImplementedError(op)
            if r.dtype not in (np.uint8, np.float32):
                raise TypeError(f"{op} returned {r.dtype}")
            r = np.ascontiguousarray(r)
            ch = 1 if r.ndim == 2 else r.shape[2]
            ptr = self.lib.ref_cv_out(out, r.shape[0], r.shape[1], (0 if r.dtype == np.uint8 else 5) + ((ch - 1) << 3))
            C.memmove(ptr, r.ctypes.data, r.nbytes)
        except Exception as e:                                             # noqa: BLE001 -- must not escape a ctypes callback
            self.errors.append(f"{op}: {e!r}")
