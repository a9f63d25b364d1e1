"""ctypes binding of libzkp_b200.so -- the same C ABI (include/zkp_b200.h) a Rust
``extern "C"`` block would bind.  There is no CPU fallback: loading fails
loudly when the CUDA library has not been built, and every call fails when no sm_100
device is present."""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libzkp_b200.so")

ZKP_OK = 0
ZKP_ERR_INVALID = -1
ZKP_ERR_CUDA = -2
ZKP_ERR_DEGREE = -3
ZKP_ERR_NOMEM = -4
ZKP_ERR_STATE = -5
ZKP_ERR_VERIFY = -6

_u64p = ctypes.POINTER(ctypes.c_uint64)
_vp = ctypes.c_void_p
_sz = ctypes.c_size_t
_int = ctypes.c_int
_uint = ctypes.c_uint

# name -> (restype, argtypes); every symbol include/zkp_b200.h declares
SIGNATURES = {
    "zkp_ctx_create": (_int, [_int, ctypes.POINTER(_vp)]),
    "zkp_ctx_destroy": (None, [_vp]),
    "zkp_last_error": (ctypes.c_char_p, [_vp]),
    "zkp_strerror": (ctypes.c_char_p, [_int]),
    "zkp_ctx_sync": (_int, [_vp]),
    "zkp_ctx_stream": (_vp, [_vp]),
    "zkp_sm_count": (_int, [_vp]),
    "zkp_launch_count": (ctypes.c_uint64, [_vp]),
    "zkp_msm_point_count": (ctypes.c_uint64, [_vp]),
    "zkp_timer_start": (_int, [_vp]),
    "zkp_timer_stop_ms": (_int, [_vp, ctypes.POINTER(ctypes.c_float)]),
    "zkp_prof_enable": (_int, [_vp, _int]),
    "zkp_prof_reset": (_int, [_vp]),
    "zkp_prof_read": (_int, [_vp, ctypes.c_char_p, ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_uint64)]),
    "zkp_buf_alloc": (_int, [_vp, _sz, ctypes.POINTER(_vp)]),
    "zkp_buf_free": (_int, [_vp, _vp]),
    "zkp_buf_len": (_sz, [_vp]),
    "zkp_buf_upload": (_int, [_vp, _vp, _sz, _vp, _sz]),
    "zkp_buf_download": (_int, [_vp, _vp, _sz, _vp, _sz]),
    "zkp_buf_wrap": (_int, [_vp, _vp, _sz, ctypes.POINTER(_vp)]),
    "zkp_buf_upload_2d": (_int, [_vp, _vp, _sz, _vp, _sz, _sz, _sz]),
    "zkp_buf_download_2d": (_int, [_vp, _vp, _sz, _vp, _sz, _sz, _sz]),
    "zkp_permute_dev": (_int, [_vp, _vp, _sz, _vp, _sz, _sz, _sz, _sz]),
    "zkp_scale_matrix_dev": (_int, [_vp, _vp, _sz, _sz, _sz, _sz, _vp, _vp, _int]),
    "zkp_buf_zero": (_int, [_vp, _vp, _sz, _sz]),
    "zkp_buf_copy": (_int, [_vp, _vp, _sz, _vp, _sz, _sz]),
    "zkp_keccak_f1600": (None, [_vp]),
    "zkp_host_alloc": (_int, [_sz, ctypes.POINTER(_vp)]),
    "zkp_host_free": (_int, [_vp]),
    "zkp_ntt": (_int, [_vp, _vp, _sz, _uint, _int, _int]),
    "zkp_ntt_dev": (_int, [_vp, _vp, _sz, _vp, _uint, _int, _int]),
    "zkp_ntt_dev_batch": (_int, [_vp, _vp, _sz, _sz, _vp, _sz, _uint, _int, _int, _uint]),
    "zkp_fft_constant": (_int, [_uint, _int, _vp]),
    "zkp_fft_elements_dev": (_int, [_vp, _uint, _vp]),
    "zkp_srs_load": (_int, [_vp, _vp, _sz, ctypes.POINTER(_vp)]),
    "zkp_srs_free": (_int, [_vp, _vp]),
    "zkp_srs_len": (_sz, [_vp]),
    "zkp_srs_generate": (_int, [_vp, _vp, _sz, ctypes.POINTER(_vp)]),
    "zkp_srs_generate_range": (_int, [_vp, _vp, _sz, _sz, ctypes.POINTER(_vp)]),
    "zkp_poly_degree_dev": (_int, [_vp, _vp, _sz, _sz, ctypes.POINTER(ctypes.c_longlong)]),
    "zkp_srs_download": (_int, [_vp, _vp, _sz, _vp, _sz]),
    "zkp_srs_trim": (_int, [_vp, _vp, _sz, ctypes.POINTER(_vp)]),
    "zkp_msm_g1": (_int, [_vp, _vp, _vp, _sz, _vp]),
    "zkp_msm_g1_dev": (_int, [_vp, _vp, _vp, _sz, _sz, _vp]),
    "zkp_commit": (_int, [_vp, _vp, _vp, _sz, _vp]),
    "zkp_commit_dev": (_int, [_vp, _vp, _vp, _sz, _sz, _vp]),
    "zkp_msm_set_window": (_int, [_vp, _uint]),
    "zkp_transcript_init": (_int, [_vp, _vp, ctypes.c_uint32]),
}



class PolyRef(ctypes.Structure):
    """``zkp_poly_ref``: buf[off .. off+len)."""
    _fields_ = [("buf", _vp), ("off", _sz), ("len", _sz)]


class QuotientArgs(ctypes.Structure):
    """``zkp_quotient_args``."""
    _fields_ = [("wires", PolyRef * 4), ("z", PolyRef), ("pi", PolyRef), ("l1", PolyRef),
                ("sel", PolyRef * 11), ("sigma", PolyRef * 4), ("linear", PolyRef),
                ("challenges", (ctypes.c_uint64 * 4) * 7), ("zh_inv", (ctypes.c_uint64 * 4) * 8),
                ("widget_mask", ctypes.c_uint32), ("coset_log_n", ctypes.c_uint32), ("coset_first", ctypes.c_uint32),
                ("sliced", ctypes.c_uint32)]


SIGNATURES.update({
    "zkp_commit_batch_dev": (_int, [_vp, _vp, ctypes.POINTER(PolyRef), _uint, _vp, ctypes.POINTER(_int)]),
    "zkp_ntt_ref_dev": (_int, [_vp, PolyRef, _vp, _sz, _uint, _int, _int]),
    "zkp_buf_fill": (_int, [_vp, _vp, _sz, _sz, _vp]),
    "zkp_poly_blind_dev": (_int, [_vp, _vp, _sz, _sz, _vp, _uint]),
    "zkp_perm_lagrange_dev": (_int, [_vp, _uint, _vp, _sz, _vp, _vp, _sz]),
    "zkp_perm_z_dev": (_int, [_vp, _sz, ctypes.POINTER(PolyRef), ctypes.POINTER(PolyRef), _vp, _vp, _vp, _vp, _sz]),
    "zkp_quotient_dev": (_int, [_vp, _uint, ctypes.POINTER(QuotientArgs), _vp, _sz]),
    "zkp_quotient_range_dev": (_int, [_vp, _uint, ctypes.POINTER(QuotientArgs), _sz, _sz, _vp, _sz]),
    "zkp_poly_eval_dev": (_int, [_vp, ctypes.POINTER(PolyRef), _uint, _vp, _vp]),
    "zkp_poly_eval2_dev": (_int, [_vp, ctypes.POINTER(PolyRef), _vp, _uint, _vp, _vp]),
    "zkp_poly_lincomb_dev": (_int, [_vp, ctypes.POINTER(PolyRef), _vp, _uint, _vp, _sz, _sz]),
    "zkp_poly_div_linear_dev": (_int, [_vp, PolyRef, _vp, _vp, _sz]),
})



class ProvingKeyDesc(ctypes.Structure):
    """``zkp_proving_key``."""
    _fields_ = [("k", _uint), ("poly", PolyRef * 15), ("eval8", PolyRef * 15), ("linear8", PolyRef),
                ("sigma_evals", PolyRef * 4), ("roots", _vp), ("zh_inv", (ctypes.c_uint64 * 4) * 8),
                ("generator", ctypes.c_uint64 * 4), ("widget_mask", ctypes.c_uint32)]


class VerifierKeyDesc(ctypes.Structure):
    """``zkp_verifier_key``."""
    _fields_ = [("k", _uint), ("constraints", ctypes.c_uint64), ("commitments", (ctypes.c_uint64 * 12) * 15)]


SIGNATURES.update({
    "zkp_prover_create": (_int, [_vp, _vp, ctypes.POINTER(ProvingKeyDesc), ctypes.POINTER(_vp)]),
    "zkp_prover_destroy": (_int, [_vp]),
    "zkp_prover_prove": (_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "zkp_prover_set_wiring": (_int, [_vp, _vp, _sz, _vp, _sz]),
    "zkp_prover_prove_witness": (_int, [_vp, _vp, _vp, _sz, _vp, _vp, _vp, _vp, _vp, _vp]),
    "zkp_transcript_append": (_int, [_vp, ctypes.c_char_p, _vp, ctypes.c_uint32]),
    "zkp_transcript_challenge": (_int, [_vp, ctypes.c_char_p, _vp, ctypes.c_uint32]),
    "zkp_linearization_scalars": (_int, [_uint, _vp, _vp, _vp]),
    "zkp_g1_compress": (_int, [_vp, _vp]),
    "zkp_fr_from_wide": (_int, [_vp, _vp]),
    # verifier glue (host code)
    "zkp_g2_generator_mul": (_int, [_vp, _vp]),
    "zkp_g1_generator_mul": (_int, [_vp, _vp]),
    "zkp_g1_decompress": (_int, [_vp, _vp]),
    "zkp_proof_decode": (_int, [_vp, _vp, _vp]),
    "zkp_pairing_check": (_int, [_vp, _vp, _sz]),
    "zkp_pairing_selftest": (_int, [_vp, _vp]),
    "zkp_kzg_batch_check": (_int, [_vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "zkp_verify": (_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz]),
    # one job over several GPUs
    "zkp_comm_unique_id": (_int, [_vp]),
    "zkp_comm_create": (_int, [_vp, _vp, _int, _int, ctypes.POINTER(_vp)]),
    "zkp_comm_destroy": (_int, [_vp]),
    "zkp_comm_rank": (_int, [_vp]),
    "zkp_comm_size": (_int, [_vp]),
    "zkp_comm_all_to_all_dev": (_int, [_vp, _vp, _sz, _vp, _sz, _sz]),
    "zkp_twiddle_transpose_dev": (_int, [_vp, _vp, _sz, _vp, _sz, _sz, _sz, _sz, _uint, _int]),
    "zkp_comm_stats": (_int, [_vp, ctypes.POINTER(ctypes.c_uint64), ctypes.POINTER(ctypes.c_uint64)]),
    "zkp_commit_batch_sharded_dev": (_int, [_vp, _vp, _vp, ctypes.POINTER(PolyRef), _uint, _vp, ctypes.POINTER(_int)]),
    "zkp_coset8_ntt_dev": (_int, [_vp, _vp, _sz, _sz, _vp, _sz, _uint, _uint, _uint]),
    "zkp_prover_create_sharded": (_int, [_vp, _vp, _vp, ctypes.POINTER(ProvingKeyDesc), ctypes.POINTER(_vp)]),
})

_lib = None


class ZkpError(RuntimeError):
    def __init__(self, code, detail=""):
        self.code = code
        msg = "zkp_b200 error %d" % code
        if _lib is not None:
            msg += " (%s)" % _lib.zkp_strerror(code).decode()
        if detail:
            msg += ": " + detail
        super().__init__(msg)


def load_library():
    """Load libzkp_b200.so and bind every exported symbol.  Raises if it is missing --
    the product has no other compute path."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "libzkp_b200.so not built (%s); run `python -c 'import __graft_entry__ as g; g.build()'`"
            % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the ABI drifted
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def _ptr(a):
    return ctypes.c_void_p(a.ctypes.data)


def as_fr_array(a):
    a = np.ascontiguousarray(a, dtype=np.uint64)
    assert a.ndim == 2 and a.shape[1] == 4, "Fr arrays are (n, 4) uint64 Montgomery limbs"
    return a


class Context:
    """One CUDA device + stream (``zkp_ctx``)."""

    def __init__(self, device=0):
        self.lib = load_library()
        h = _vp()
        rc = self.lib.zkp_ctx_create(device, ctypes.byref(h))
        if rc:
            raise ZkpError(rc, "zkp_ctx_create(device=%d): no usable sm_100 GPU" % device)
        self.h = h
        self.device = device

    def check(self, rc):
        if rc:
            raise ZkpError(rc, self.lib.zkp_last_error(self.h).decode() if rc == ZKP_ERR_CUDA else "")

    def close(self):
        if getattr(self, "h", None):
            self.lib.zkp_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def sync(self):
        self.check(self.lib.zkp_ctx_sync(self.h))

    @property
    def stream(self):
        return self.lib.zkp_ctx_stream(self.h)

    @property
    def sm_count(self):
        return self.lib.zkp_sm_count(self.h)

    @property
    def launches(self):
        return int(self.lib.zkp_launch_count(self.h))

    @property
    def msm_points(self):
        return int(self.lib.zkp_msm_point_count(self.h))

    def timer_start(self):
        self.check(self.lib.zkp_timer_start(self.h))

    def timer_stop_ms(self):
        ms = ctypes.c_float()
        self.check(self.lib.zkp_timer_stop_ms(self.h, ctypes.byref(ms)))
        return ms.value

    def prof_enable(self, on=True):
        self.check(self.lib.zkp_prof_enable(self.h, int(on)))

    def prof_reset(self):
        self.check(self.lib.zkp_prof_reset(self.h))

    def prof_read(self, name):
        ms = ctypes.c_float()
        cnt = ctypes.c_uint64()
        self.check(self.lib.zkp_prof_read(self.h, name.encode(), ctypes.byref(ms), ctypes.byref(cnt)))
        return ms.value, int(cnt.value)

    # ---- buffers
    def alloc(self, n):
        return DeviceBuffer(self, n)

    def upload(self, arr):
        arr = as_fr_array(arr)
        b = DeviceBuffer(self, arr.shape[0])
        b.upload(arr)
        return b

    # ---- NTT on host vectors
    def ntt(self, data, k, inverse=False, coset=False):
        data = as_fr_array(data)
        n = 1 << k
        m = data.shape[0]
        assert m <= n
        buf = np.zeros((n, 4), dtype=np.uint64)
        buf[:m] = data
        self.check(self.lib.zkp_ntt(self.h, _ptr(buf), m, k, int(inverse), int(coset)))
        return buf

    def ntt_dev(self, src, len_in, dst, k, inverse=False, coset=False):
        """src / dst: DeviceBuffer or BufferView."""
        sb, so = (src.buf, src.off) if isinstance(src, BufferView) else (src, 0)
        db, do = (dst.buf, dst.off) if isinstance(dst, BufferView) else (dst, 0)
        self.check(self.lib.zkp_ntt_ref_dev(self.h, PolyRef(sb.h, so, len_in), db.h, do, k, int(inverse),
                                            int(coset)))

    def ntt_dev_batch(self, src, in_stride, len_in, dst, out_stride, k, inverse, coset, batch):
        """``batch`` same-shape transforms; src / dst must be whole DeviceBuffers (offset 0)."""
        self.check(self.lib.zkp_ntt_dev_batch(self.h, src.h, in_stride, len_in, dst.h, out_stride, k,
                                              int(inverse), int(coset), batch))

    def wrap(self, device_ptr, n):
        """DeviceBuffer over caller-owned device memory (``zkp_buf_wrap``)."""
        return DeviceBuffer(self, n, device_ptr=device_ptr)

    def permute(self, src, src_off, dst, dst_off, A, B, w=1):
        self.check(self.lib.zkp_permute_dev(self.h, src.h, src_off, dst.h, dst_off, A, B, w))

    def scale_matrix(self, buf, off, rows, cols, a0, base1, base2, mode):
        b1 = np.ascontiguousarray(base1, dtype=np.uint64).reshape(4)
        b2 = np.ascontiguousarray(base2, dtype=np.uint64).reshape(4)
        self.check(self.lib.zkp_scale_matrix_dev(self.h, buf.h, off, rows, cols, a0, _ptr(b1), _ptr(b2), mode))

    def fft_elements(self, k):
        b = DeviceBuffer(self, 1 << k)
        self.check(self.lib.zkp_fft_elements_dev(self.h, k, b.h))
        return b

    # ---- SRS / MSM
    def srs_load(self, xy):
        xy = np.ascontiguousarray(xy, dtype=np.uint64).reshape(-1, 12)
        return Srs(self, xy=xy)

    def srs_generate(self, tau_mont, n, first=0):
        return Srs(self, tau=np.ascontiguousarray(tau_mont, dtype=np.uint64).reshape(4), n=n, first=first)

    def poly_degree(self, buf, off=0, n=None):
        """Highest non-zero index of buf[off .. off+n), -1 if all zero."""
        n = buf.n - off if n is None else n
        top = ctypes.c_longlong()
        self.check(self.lib.zkp_poly_degree_dev(self.h, buf.h, off, n, ctypes.byref(top)))
        return int(top.value)

    def msm(self, srs, scalars):
        scalars = as_fr_array(scalars)
        out = np.zeros(12, dtype=np.uint64)
        self.check(self.lib.zkp_msm_g1(self.h, srs.h, _ptr(scalars), scalars.shape[0], _ptr(out)))
        return out

    def msm_dev(self, srs, buf, off=0, n=None):
        n = buf.n - off if n is None else n
        out = np.zeros(12, dtype=np.uint64)
        self.check(self.lib.zkp_msm_g1_dev(self.h, srs.h, buf.h, off, n, _ptr(out)))
        return out

    def commit(self, srs, coeffs):
        coeffs = as_fr_array(coeffs)
        out = np.zeros(12, dtype=np.uint64)
        rc = self.lib.zkp_commit(self.h, srs.h, _ptr(coeffs), coeffs.shape[0], _ptr(out))
        self.check(rc)
        return out

    def commit_dev(self, srs, buf, off=0, n=None):
        n = buf.n - off if n is None else n
        out = np.zeros(12, dtype=np.uint64)
        self.check(self.lib.zkp_commit_dev(self.h, srs.h, buf.h, off, n, _ptr(out)))
        return out

    def commit_batch_dev(self, srs, refs):
        """-> ((count, 12) uint64 affine Montgomery, [status per polynomial])."""
        arr = (PolyRef * len(refs))(*refs)
        out = np.zeros((len(refs), 12), dtype=np.uint64)
        st = (_int * len(refs))()
        self.check(self.lib.zkp_commit_batch_dev(self.h, srs.h, arr, len(refs), _ptr(out), st))
        return out, list(st)

    def commit_batch_sharded_dev(self, comm, srs, refs):
        """``commit_batch_dev`` over every rank of ``comm`` (collective): SRS ranges + partial-sum gather."""
        arr = (PolyRef * len(refs))(*refs)
        out = np.zeros((len(refs), 12), dtype=np.uint64)
        st = (_int * len(refs))()
        self.check(self.lib.zkp_commit_batch_sharded_dev(self.h, comm.h if comm is not None else None, srs.h, arr,
                                                         len(refs), _ptr(out), st))
        return out, list(st)

    def twiddle_transpose(self, src, src_off, dst, dst_off, rows, cols, a0, k, inverse):
        """dst[b][a] = src[a][b] * w_N^(+-(a0 + a) b): the four-step twiddle fused into the transpose."""
        sb, so = _base(src)
        db, do = _base(dst)
        self.check(self.lib.zkp_twiddle_transpose_dev(self.h, sb.h, so + src_off, db.h, do + dst_off, rows, cols, a0, k,
                                                      int(inverse)))

    def coset8_ntt(self, src, src_off, len_in, dst, dst_off, k, first, count):
        """dst[(u - first) n + m] = p(g w_8n^u w_n^m): the polynomial's values on whole cosets of the 8n domain."""
        sb, so = _base(src)
        db, do = _base(dst)
        self.check(self.lib.zkp_coset8_ntt_dev(self.h, sb.h, so + src_off, len_in, db.h, do + dst_off, k, first, count))

    def set_msm_window(self, c):
        self.check(self.lib.zkp_msm_set_window(self.h, c))

    # ---- prover rounds (device-resident)
    @staticmethod
    def ref(buf, off=0, n=None):
        n = (buf.n - off) if n is None else n
        base, boff = _base(buf)
        return PolyRef(base.h, boff + off, n)

    def fill(self, buf, off, n, value):
        v = np.ascontiguousarray(value, dtype=np.uint64).reshape(4)
        base, boff = _base(buf)
        self.check(self.lib.zkp_buf_fill(self.h, base.h, boff + off, n, _ptr(v)))

    def poly_blind(self, buf, off, n, blinders):
        b = as_fr_array(blinders)
        base, boff = _base(buf)
        self.check(self.lib.zkp_poly_blind_dev(self.h, base.h, boff + off, n, _ptr(b), b.shape[0]))

    def perm_lagrange(self, k, enc, roots, out, out_off=0):
        enc = np.ascontiguousarray(enc, dtype=np.uint32)
        self.check(self.lib.zkp_perm_lagrange_dev(self.h, k, _ptr(enc), enc.shape[0], roots.h, out.h, out_off))

    def perm_z(self, n, wires, sigmas, roots, beta, gamma, out, out_off=0):
        w = (PolyRef * 4)(*wires)
        s = (PolyRef * 4)(*sigmas)
        b = np.ascontiguousarray(beta, dtype=np.uint64).reshape(4)
        g = np.ascontiguousarray(gamma, dtype=np.uint64).reshape(4)
        self.check(self.lib.zkp_perm_z_dev(self.h, n, w, s, roots.h, _ptr(b), _ptr(g), out.h, out_off))

    def quotient(self, k8, args, out, out_off=0, first=None, count=None):
        if first is None:
            self.check(self.lib.zkp_quotient_dev(self.h, k8, ctypes.byref(args), out.h, out_off))
        else:
            self.check(self.lib.zkp_quotient_range_dev(self.h, k8, ctypes.byref(args), first, count, out.h, out_off))

    def poly_eval(self, refs, point):
        arr = (PolyRef * len(refs))(*refs)
        pt = np.ascontiguousarray(point, dtype=np.uint64).reshape(4)
        out = np.zeros((len(refs), 4), dtype=np.uint64)
        self.check(self.lib.zkp_poly_eval_dev(self.h, arr, len(refs), _ptr(pt), _ptr(out)))
        return out

    def poly_eval2(self, refs, which, points):
        """Polynomial i at points[which[i]] (two points, one launch)."""
        arr = (PolyRef * len(refs))(*refs)
        w = np.ascontiguousarray(which, dtype=np.uint8)
        pts = np.ascontiguousarray(points, dtype=np.uint64).reshape(8)
        out = np.zeros((len(refs), 4), dtype=np.uint64)
        self.check(self.lib.zkp_poly_eval2_dev(self.h, arr, _ptr(w), len(refs), _ptr(pts), _ptr(out)))
        return out

    def poly_lincomb(self, refs, scalars, out, out_off, out_len):
        arr = (PolyRef * len(refs))(*refs)
        sc = as_fr_array(scalars)
        assert sc.shape[0] == len(refs)
        self.check(self.lib.zkp_poly_lincomb_dev(self.h, arr, _ptr(sc), len(refs), out.h, out_off, out_len))

    def poly_div_linear(self, ref, point, out, out_off=0):
        pt = np.ascontiguousarray(point, dtype=np.uint64).reshape(4)
        self.check(self.lib.zkp_poly_div_linear_dev(self.h, ref, _ptr(pt), out.h, out_off))


_PINNED = {}   # data address -> (ctypes buffer, raw pointer); keeps the mapping alive


def pinned_empty(shape, dtype=np.uint64):
    """numpy array backed by page-locked host memory (``zkp_host_alloc``).  Release with
    ``pinned_free``; otherwise it lives until process exit."""
    lib = load_library()
    count = int(np.prod(shape))
    nbytes = count * np.dtype(dtype).itemsize
    p = _vp()
    rc = lib.zkp_host_alloc(nbytes, ctypes.byref(p))
    if rc:
        raise ZkpError(rc, "zkp_host_alloc(%d bytes)" % nbytes)
    buf = (ctypes.c_char * max(nbytes, 1)).from_address(p.value)
    arr = np.frombuffer(buf, dtype=dtype, count=count).reshape(shape)
    _PINNED[arr.ctypes.data] = (buf, p)
    return arr


def pinned_free(arr):
    ent = _PINNED.pop(arr.ctypes.data, None)
    if ent is not None:
        load_library().zkp_host_free(ent[1])


def fft_constant(k, kind):
    lib = load_library()
    out = np.zeros(4, dtype=np.uint64)
    rc = lib.zkp_fft_constant(k, kind, _ptr(out))
    if rc:
        raise ZkpError(rc)
    return out


class BufferView:
    """buf[off .. off+n): a polynomial living inside a larger device buffer."""

    def __init__(self, buf, off=0, n=None):
        if n is None:
            n = buf.n - off
        if isinstance(buf, BufferView):
            buf, off = buf.buf, buf.off + off
        self.buf, self.off, self.n = buf, off, n

    def upload(self, arr, off=0):
        self.buf.upload(arr, self.off + off)

    def download(self, off=0, n=None):
        return self.buf.download(self.off + off, (self.n - off) if n is None else n)

    def zero(self, off=0, n=None):
        self.buf.zero(self.off + off, (self.n - off) if n is None else n)


def _base(b):
    """(DeviceBuffer, element offset) of a DeviceBuffer or BufferView."""
    return (b.buf, b.off) if isinstance(b, BufferView) else (b, 0)


class DeviceBuffer:
    """Device-resident Fr vector (``zkp_buf``)."""

    def __init__(self, ctx, n, device_ptr=None):
        self.ctx = ctx
        self.n = n
        h = _vp()
        if device_ptr is None:
            ctx.check(ctx.lib.zkp_buf_alloc(ctx.h, n, ctypes.byref(h)))
        else:
            ctx.check(ctx.lib.zkp_buf_wrap(ctx.h, _vp(device_ptr), n, ctypes.byref(h)))
        self.h = h

    def upload(self, arr, off=0):
        arr = as_fr_array(arr)
        self.ctx.check(self.ctx.lib.zkp_buf_upload(self.ctx.h, self.h, off, _ptr(arr), arr.shape[0]))

    def download(self, off=0, n=None):
        n = self.n - off if n is None else n
        out = np.empty((n, 4), dtype=np.uint64)
        self.ctx.check(self.ctx.lib.zkp_buf_download(self.ctx.h, self.h, off, _ptr(out), n))
        return out

    def zero(self, off=0, n=None):
        n = self.n - off if n is None else n
        self.ctx.check(self.ctx.lib.zkp_buf_zero(self.ctx.h, self.h, off, n))

    def copy_from(self, src, n, dst_off=0, src_off=0):
        self.ctx.check(self.ctx.lib.zkp_buf_copy(self.ctx.h, self.h, dst_off, src.h, src_off, n))

    def free(self):
        if self.h and self.ctx.h:
            self.ctx.lib.zkp_buf_free(self.ctx.h, self.h)
        self.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Srs:
    """Device-resident SRS powers (``zkp_srs``)."""

    def __init__(self, ctx, xy=None, tau=None, n=None, first=0):
        self.ctx = ctx
        h = _vp()
        if xy is not None:
            ctx.check(ctx.lib.zkp_srs_load(ctx.h, _ptr(xy), xy.shape[0], ctypes.byref(h)))
            self.n = xy.shape[0]
        else:
            ctx.check(ctx.lib.zkp_srs_generate_range(ctx.h, _ptr(tau), first, n, ctypes.byref(h)))
            self.n = n
        self.h = h

    def download(self, off=0, n=None):
        n = self.n - off if n is None else n
        out = np.empty((n, 12), dtype=np.uint64)
        self.ctx.check(self.ctx.lib.zkp_srs_download(self.ctx.h, self.h, off, _ptr(out), n))
        return out

    def trim(self, keep):
        """The first ``keep`` powers as a new SRS, without leaving the device (``zkp_srs_trim``)."""
        t = object.__new__(Srs)
        t.ctx, t.n = self.ctx, keep
        h = _vp()
        self.ctx.check(self.ctx.lib.zkp_srs_trim(self.ctx.h, self.h, keep, ctypes.byref(h)))
        t.h = h
        return t

    def free(self):
        if self.h and self.ctx.h:
            self.ctx.lib.zkp_srs_free(self.ctx.h, self.h)
        self.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class NativeProver:
    """``zkp_prover``: the round driver of ``create_proof`` in native code (csrc/create_proof.cu)."""

    def __init__(self, ctx, srs, desc, keepalive, comm=None, sharded=False):
        self.ctx = ctx
        self._keep = (srs, desc, keepalive, comm)   # the key's buffers, the SRS and the communicator must outlive the handle
        h = _vp()
        if sharded:
            ctx.check(ctx.lib.zkp_prover_create_sharded(ctx.h, comm.h if comm is not None else None, srs.h,
                                                        ctypes.byref(desc), ctypes.byref(h)))
        else:
            ctx.check(ctx.lib.zkp_prover_create(ctx.h, srs.h, ctypes.byref(desc), ctypes.byref(h)))
        self.h = h
        self._comms = np.zeros((11, 12), dtype=np.uint64)
        self._evals = np.zeros((16, 4), dtype=np.uint64)
        self._bytes = np.zeros(1040, dtype=np.uint8)

    def prove(self, transcript_state, wires_host, wires_dev, pi_host, pi_dev, blinders_mont):
        """-> (rc, commitments (11, 12), evaluations (16, 4), proof bytes).  Exactly one of
        wires_host / wires_dev and of pi_host / pi_dev is given."""
        st = np.frombuffer(bytes(transcript_state), dtype=np.uint8)
        assert st.shape[0] == 203
        bl = as_fr_array(blinders_mont)
        assert bl.shape[0] == 11
        rc = self.ctx.lib.zkp_prover_prove(
            self.h, _ptr(st),
            _ptr(wires_host) if wires_host is not None else None, wires_dev.h if wires_dev is not None else None,
            _ptr(pi_host) if pi_host is not None else None, pi_dev.h if pi_dev is not None else None,
            _ptr(bl), _ptr(self._comms), _ptr(self._evals), _ptr(self._bytes), None)
        return rc, self._comms, self._evals, self._bytes

    def set_wiring(self, wire_idx, pi_idx):
        """wire_idx: (4, m) witness indices per gate; pi_idx: gate positions of the public inputs."""
        w = np.ascontiguousarray(wire_idx, dtype=np.uint32)
        p = np.ascontiguousarray(pi_idx, dtype=np.uint32)
        assert w.ndim == 2 and w.shape[0] == 4
        self.ctx.check(self.ctx.lib.zkp_prover_set_wiring(self.h, _ptr(w), w.shape[1], _ptr(p) if p.size else None,
                                                          p.size))

    def prove_witness(self, transcript_state, witness_mont, pi_values_mont, blinders_mont):
        """As ``prove`` with the wire gather done on the device from the witness values."""
        st = np.frombuffer(bytes(transcript_state), dtype=np.uint8)
        assert st.shape[0] == 203
        bl = as_fr_array(blinders_mont)
        wv = as_fr_array(witness_mont)
        pv = as_fr_array(pi_values_mont) if len(pi_values_mont) else None
        rc = self.ctx.lib.zkp_prover_prove_witness(
            self.h, _ptr(st), _ptr(wv), wv.shape[0], _ptr(pv) if pv is not None else None,
            _ptr(bl), _ptr(self._comms), _ptr(self._evals), _ptr(self._bytes), None)
        return rc, self._comms, self._evals, self._bytes

    def close(self):
        if getattr(self, "h", None) and self.ctx.h:
            self.ctx.lib.zkp_prover_destroy(self.h)
        self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class NativeComm:
    """``zkp_comm``: this rank's link to the GPUs that share one job (NCCL inside libzkp_b200.so).
    ``world == 1`` needs no NCCL: the sharded code path on one GPU."""

    def __init__(self, ctx, rank=0, world=1, unique_id=None):
        self.ctx, self.rank, self.world = ctx, rank, world
        h = _vp()
        if world > 1:
            assert unique_id is not None and len(unique_id) == 256
            uid = np.frombuffer(bytes(unique_id), dtype=np.uint8).copy()
            ctx.check(ctx.lib.zkp_comm_create(ctx.h, _ptr(uid), rank, world, ctypes.byref(h)))
        else:
            ctx.check(ctx.lib.zkp_comm_create(ctx.h, None, 0, 1, ctypes.byref(h)))
        self.h = h

    @staticmethod
    def unique_id():
        out = np.zeros(256, dtype=np.uint8)
        rc = load_library().zkp_comm_unique_id(_ptr(out))
        if rc:
            raise ZkpError(rc, "NCCL unavailable")
        return bytes(out)

    @classmethod
    def from_torch_distributed(cls, ctx):
        """One rank per process under torchrun: rank 0 makes the id, torch.distributed carries its 256 bytes."""
        import torch.distributed as dist
        rank, world = dist.get_rank(), dist.get_world_size()
        box = [cls.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        return cls(ctx, rank, world, box[0])

    def all_to_all(self, src, dst, count):
        """Block p of ``src`` -> rank p, block r of ``dst`` <- rank r (``count`` Fr each), on the context's stream."""
        sb, so = _base(src)
        db, do = _base(dst)
        self.ctx.check(self.ctx.lib.zkp_comm_all_to_all_dev(self.h, sb.h, so, db.h, do, count))

    def stats(self):
        c, b = ctypes.c_uint64(), ctypes.c_uint64()
        self.ctx.check(self.ctx.lib.zkp_comm_stats(self.h, ctypes.byref(c), ctypes.byref(b)))
        return int(c.value), int(b.value)

    def close(self):
        if getattr(self, "h", None) and self.ctx.h:
            self.ctx.lib.zkp_comm_destroy(self.h)
        self.h = None
