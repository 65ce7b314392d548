"""Engine: one GPU context of the C-ABI library, holding the device-resident matrix shard.

This is the host-side plumbing the API modules (normalize, hvg, dr, clustering) share.  All
numerics run in libadobo_b200.so; nothing here computes on the CPU.
"""
import ctypes as C
import os

import numpy as np
import scipy.sparse as sp

from . import _lib
from ._lib import check

LOG_NONE, LOG_2, LOG_E, LOG_10 = 0, 1, 2, 3


def as_csr(m):
    """scipy CSR with int64 indptr / int32 indices, sorted column ids (layout of the C-ABI)."""
    if not isinstance(m, sp.csr_matrix):             # an existing CSR object keeps its cached has_sorted_indices flag
        m = sp.csr_matrix(m)                         # (the check is a pass over all non-zeros: 35 ms at 70 M)
    if not m.has_sorted_indices:
        m = m.sorted_indices()
    return m


class LanczosResult:
    """Same attributes as adobo.irlbpy.irlb.LanczosResult (irlb.py:261-265)."""

    def __init__(self, **kw):
        for k, v in kw.items():
            setattr(self, k, v)


class Engine:
    def __init__(self, device=None):
        self.lib = _lib.load()
        if device is None:
            device = int(os.environ.get('ADOBO_B200_DEVICE', os.environ.get('LOCAL_RANK', '0')))
        self.device = device
        h = C.c_void_p()
        check(self.lib.adobo_ctx_create(int(device), C.byref(h)))
        self.h = h
        self.world, self.rank = 1, 0
        self.n = self.g = self.nnz = 0
        self.resident_key = None
        self.counts_key = None                           # data.ResidentKey of the count frame whose matrix is resident
        self._pattern = None
        self._lazy = []                                  # weakrefs to data.DeviceCsrFrame objects without a host copy

    def _evict(self):
        """Before the resident COUNTS are replaced: every normalized frame whose values were left on the device
        (data.DeviceCsrFrame: re-derived from the resident counts on demand) is brought to the host first, so that
        no frame ever loses its data."""
        lazy, self._lazy = self._lazy, []
        for ref in lazy:
            lz = ref()
            if lz is not None:
                lz.materialize()
        self.counts_key = None

    def close(self):
        if getattr(self, 'h', None):
            try:
                self._evict()                            # frames whose values live here only get their host copy first
            except Exception:
                pass
            self.lib.adobo_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- multi-GPU ----
    def init_comm_from_torch(self):
        """Joins the contexts of a torchrun job: rank 0 makes the NCCL id, torch.distributed
        (any backend) broadcasts it."""
        import torch.distributed as dist
        world, rank = dist.get_world_size(), dist.get_rank()
        ident = np.zeros(128, dtype=np.uint8)
        if rank == 0:
            check(self.lib.adobo_comm_unique_id(ident))
        obj = [ident.tobytes()]
        dist.broadcast_object_list(obj, src=0)
        ident = np.frombuffer(obj[0], dtype=np.uint8).copy()
        check(self.lib.adobo_comm_init(self.h, world, rank, ident))
        self.world, self.rank = world, rank

    # ---- stopwatch ----
    def record(self, slot):
        check(self.lib.adobo_event_record(self.h, slot))

    def elapsed_ms(self, a, b):
        ms = C.c_float()
        check(self.lib.adobo_event_elapsed_ms(self.h, a, b, C.byref(ms)))
        return ms.value

    def sync(self):
        check(self.lib.adobo_sync(self.h))

    def l2_flush(self):
        check(self.lib.adobo_l2_flush(self.h))

    def launch_count(self):
        return int(self.lib.adobo_launch_count(self.h))

    def profile(self, on):
        check(self.lib.adobo_profile(self.h, 1 if on else 0))

    def profile_read(self):
        n = 64
        names = C.create_string_buffer(32 * n)
        ms = np.zeros(n, dtype=np.float32)
        cnt = np.zeros(n, dtype=np.int64)
        work = np.zeros(n, dtype=np.float64)
        k = C.c_int()
        check(self.lib.adobo_profile_read(self.h, n, names, ms, cnt, work, C.byref(k)))
        out = {}
        for i in range(k.value):
            nm = names.raw[32 * i:32 * i + 32].split(b'\0')[0].decode()
            out[nm] = dict(ms=float(ms[i]), launches=int(cnt[i]), work=float(work[i]))
        return out

    def pin(self, arr):
        """Page-locks a numpy array in place (cudaHostRegister)."""
        check(self.lib.adobo_host_register(arr.ctypes.data, arr.nbytes))

    def unpin(self, arr):
        check(self.lib.adobo_host_unregister(arr.ctypes.data))

    # ---- data in ----
    def upload_counts(self, counts, key=None):
        self._evict()
        m = as_csr(counts)
        indptr = np.ascontiguousarray(m.indptr, dtype=np.int64)
        indices = np.ascontiguousarray(m.indices, dtype=np.int32)
        data = np.ascontiguousarray(m.data)
        if data.dtype != np.int32 and data.size and (data.min() < -2 ** 31 or data.max() >= 2 ** 31):
            raise Exception('counts do not fit int32')
        data = data.astype(np.int32, copy=False)
        check(self.lib.adobo_upload_counts(self.h, m.shape[0], m.shape[1], indptr, indices, np.ascontiguousarray(data)))
        self.n, self.g, self.nnz = m.shape[0], m.shape[1], int(indptr[-1])
        self._pattern = (indptr, indices)
        self.resident_key = None
        self.counts_key = key

    def count_stats(self):
        """data.py:75-84 from the resident counts: (total_reads, detected_genes) per cell, expressed per gene."""
        tot = np.empty(self.n, dtype=np.int64)
        det = np.empty(self.n, dtype=np.int64)
        expr = np.empty(self.g, dtype=np.int64)
        check(self.lib.adobo_count_stats(self.h, tot, det, expr))
        return tot, det, expr

    def upload_normalized(self, norm, key=None):
        self._evict()
        m = as_csr(norm)
        indptr = np.ascontiguousarray(m.indptr, dtype=np.int64)
        indices = np.ascontiguousarray(m.indices, dtype=np.int32)
        data = np.ascontiguousarray(m.data, dtype=np.float64)
        check(self.lib.adobo_upload_normalized(self.h, m.shape[0], m.shape[1], indptr, indices, data))
        self.n, self.g, self.nnz = m.shape[0], m.shape[1], int(indptr[-1])
        self._pattern = (indptr, indices)
        self.resident_key = key

    def synth_counts(self, model, row_begin, n_rows):
        """Generates cells [row_begin, row_begin + n_rows) of the seeded synthetic matrix of `model`
        (adobo_b200.synth.CountModel) ON the device, resident like after upload_counts.  Returns nnz."""
        self._evict()
        table = np.ascontiguousarray(model.lam_table(), dtype=np.float32)
        nnz = C.c_int64()
        check(self.lib.adobo_synth_counts(self.h, int(n_rows), int(model.n_genes), int(row_begin), int(model.seed) & (2**64 - 1),
                                          table, int(table.shape[0]), C.byref(nnz)))
        self.n, self.g, self.nnz = int(n_rows), int(model.n_genes), int(nnz.value)
        self._pattern = None
        self.resident_key = None
        return self.nnz

    def download_counts(self):
        """The resident count shard as scipy CSR (int64 indptr, int32 indices / counts) -- for the CPU checkers."""
        indptr = np.empty(self.n + 1, dtype=np.int64)
        indices = np.empty(self.nnz, dtype=np.int32)
        data = np.empty(self.nnz, dtype=np.int32)
        check(self.lib.adobo_download_counts(self.h, indptr, indices, data))
        return sp.csr_matrix((data, indices, indptr), shape=(self.n, self.g))

    # ---- stages ----
    def normalize(self, scaling_factor=10000, log_mode=LOG_2, small_const=1.0, want_values=True):
        lib = np.empty(self.n, dtype=np.int64)
        vals = np.empty(self.nnz, dtype=np.float64) if want_values else None
        check(self.lib.adobo_normalize(self.h, float(scaling_factor), int(log_mode), float(small_const), lib, vals))
        if not want_values:
            return None, lib
        if self._pattern is None:                        # matrix generated on the device
            m = self.download_counts()
            self._pattern = (np.ascontiguousarray(m.indptr, dtype=np.int64), np.ascontiguousarray(m.indices, dtype=np.int32))
        indptr, indices = self._pattern
        out = sp.csr_matrix((vals, indices, indptr), shape=(self.n, self.g))
        out.has_sorted_indices = True                    # the pattern that was uploaded (as_csr sorts): no re-check per use
        return out, lib

    def gene_moments(self):
        mean = np.empty(self.g)
        var = np.empty(self.g)
        check(self.lib.adobo_gene_moments(self.h, mean, var))
        return mean, var

    def hvg_seurat(self, ngenes=1000, num_bins=20):
        idx = np.empty(max(int(ngenes), 1), dtype=np.int32)
        z = np.empty(self.g)
        m = C.c_int32()
        check(self.lib.adobo_hvg_seurat(self.h, int(ngenes), int(num_bins), idx, C.byref(m), z))
        return idx[:m.value].astype(np.int64), z

    def irlb(self, genes, ncomp=75, tol=0.0001, maxit=1000, v0=None, seed=None, scale=True, var_weigh=True):
        genes = np.ascontiguousarray(genes, dtype=np.int32)
        h = genes.shape[0]
        if v0 is None:
            np.random.seed(seed)                 # irlb.py:151-152 (legacy MT19937 stream)
            v0 = np.random.randn(h)
        v0 = np.ascontiguousarray(v0, dtype=np.float64)
        comp = np.empty((self.n, ncomp))
        contr = np.empty((h, ncomp))
        s = np.empty(ncomp)
        steps, nmult = C.c_int32(), C.c_int32()
        check(self.lib.adobo_irlb(self.h, genes, h, int(ncomp), float(tol), int(maxit), v0, 1 if scale else 0,
                                  1 if var_weigh else 0, comp, s, contr, C.byref(steps), C.byref(nmult)))
        return dict(comp=comp, contr=contr, s=s, steps=steps.value, nmult=nmult.value, genes=np.sort(genes))

    def gather_subset(self, genes, scale=True):
        """Leaves the column subset `genes` of the resident normalized matrix (and its mu / sigma) on the device."""
        genes = np.ascontiguousarray(genes, dtype=np.int32)
        check(self.lib.adobo_gather_subset(self.h, genes, genes.shape[0], 1 if scale else 0))
        self._subset_h = int(genes.shape[0])

    def irlb_permuted(self, perm_cols, perm_seed, ncomp=75, tol=0.0001, maxit=1000, v0=None, seed=None, var_weigh=True):
        """One dr.jackstraw permutation on the device (adobo_irlb_permuted): returns contr = V (h x ncomp)."""
        h = self._subset_h
        perm_cols = np.ascontiguousarray(perm_cols, dtype=np.int32)
        if v0 is None:
            np.random.seed(seed)                 # irlb.py:151-152
            v0 = np.random.randn(h)
        v0 = np.ascontiguousarray(v0, dtype=np.float64)
        contr = np.empty((h, ncomp))
        steps, nmult = C.c_int32(), C.c_int32()
        check(self.lib.adobo_irlb_permuted(self.h, perm_cols, perm_cols.shape[0], int(perm_seed) & (2 ** 64 - 1), int(ncomp),
                                           float(tol), int(maxit), v0, 1 if var_weigh else 0, contr, C.byref(steps),
                                           C.byref(nmult)))
        return contr

    def lanczos_csr(self, A, nval, tol=0.0001, maxit=50, seed=None, v0=None):
        m = as_csr(A)
        indptr = np.ascontiguousarray(m.indptr, dtype=np.int64)
        indices = np.ascontiguousarray(m.indices, dtype=np.int32)
        data = np.ascontiguousarray(m.data, dtype=np.float64)
        rows, cols = m.shape
        if v0 is None:
            np.random.seed(seed)
            v0 = np.random.randn(cols)
        v0 = np.ascontiguousarray(v0, dtype=np.float64)
        U = np.empty((rows, nval))
        V = np.empty((cols, nval))
        s = np.empty(nval)
        steps, nmult = C.c_int32(), C.c_int32()
        check(self.lib.adobo_lanczos_csr(self.h, rows, cols, indptr, indices, data, int(nval), float(tol), int(maxit),
                                         v0, U, s, V, C.byref(steps), C.byref(nmult)))
        return LanczosResult(U=U, s=s, V=V, steps=steps.value, nmult=nmult.value)

    def knn(self, comp=None, k=10, n_local=None, fetch=True, metric=0):
        if comp is not None:
            comp = np.ascontiguousarray(comp, dtype=np.float64)
            n, p = comp.shape
        else:
            n, p = (self.n if n_local is None else n_local), 0
        out = np.empty((n, k), dtype=np.int64) if fetch else None
        check(self.lib.adobo_knn_metric(self.h, comp, n, p, int(k), int(metric), out))
        return out

    def upload_embedding(self, comp):
        """This rank's rows of an embedding (n x p float64) left on the device for knn(None, k)."""
        comp = np.ascontiguousarray(comp, dtype=np.float64)
        check(self.lib.adobo_upload_embedding(self.h, comp, comp.shape[0], comp.shape[1]))
        self.n = comp.shape[0]
        self.resident_key = None

    def snn(self, nn_idx=None, k=10, prune_snn=0.067, fetch=True):
        ne, pr, tot = C.c_int64(), C.c_int64(), C.c_int64()
        if nn_idx is not None:
            nn_idx = np.ascontiguousarray(nn_idx, dtype=np.int64)
            n = nn_idx.shape[0]
            if nn_idx.shape[1] != k:
                raise Exception('snn: nn_idx has %d columns but k=%d (run knn with the same k)' % (nn_idx.shape[1], k))
            check(self.lib.adobo_snn(self.h, nn_idx, n, int(k), float(prune_snn), C.byref(ne), C.byref(pr),
                                     C.byref(tot)))
        else:
            check(self.lib.adobo_snn(self.h, None, 0, int(k), float(prune_snn), C.byref(ne), C.byref(pr), C.byref(tot)))
        if not fetch:
            return None, pr.value, tot.value, ne.value
        src = np.empty(max(ne.value, 1), dtype=np.int32)
        dst = np.empty(max(ne.value, 1), dtype=np.int32)
        check(self.lib.adobo_snn_fetch(self.h, src, dst))
        edges = np.stack([src[:ne.value], dst[:ne.value]], axis=1)
        return edges, pr.value, tot.value, ne.value

    def embed_path(self, v0, scaling_factor=10000, ngenes=1000, ncomp=75, k=10, prune_snn=0.067):
        """normalize -> seurat -> irlb pca -> knn -> snn, device resident (no host copies)."""
        steps, nmult, ne = C.c_int32(), C.c_int32(), C.c_int64()
        v0 = np.ascontiguousarray(v0, dtype=np.float64)
        check(self.lib.adobo_embed_path(self.h, float(scaling_factor), int(ngenes), int(ncomp), v0, int(k),
                                        float(prune_snn), C.byref(steps), C.byref(nmult), C.byref(ne)))
        return dict(steps=steps.value, nmult=nmult.value, n_edges=ne.value)

    def result_sizes(self):
        """What the context holds: sizes for the fetch buffers (never the caller's own numbers)."""
        out = np.zeros(8, dtype=np.int64)
        check(self.lib.adobo_result_sizes(self.h, out))
        return dict(zip(('n', 'ncomp', 'h', 'k', 'knn_n', 'n_edges', 'nnz', 'g'), (int(v) for v in out)))

    def fetch_pca(self, h=None, ncomp=None, out=None):
        """PCA results to the host.  out: dict(comp, contr, s, genes) of caller-owned arrays (e.g. page-locked with
        pin(), reused from pass to pass) -- the C-ABI's caller-allocated outputs; default: fresh arrays."""
        sz = self.result_sizes()
        if (h is not None and h != sz['h']) or (ncomp is not None and ncomp != sz['ncomp']):
            raise Exception('fetch_pca: the device holds %d genes x %d components, not %s x %s' % (sz['h'], sz['ncomp'], h, ncomp))
        h, ncomp = sz['h'], sz['ncomp']
        if out is not None:
            comp, contr, s, genes = out['comp'], out['contr'], out['s'], out['genes']
            if comp.shape != (sz['n'], ncomp) or contr.shape != (h, ncomp) or s.shape != (ncomp,) or genes.shape != (h,):
                raise Exception('fetch_pca: out arrays do not have the shapes of the results on the device')
        else:
            comp = np.empty((sz['n'], ncomp))
            contr = np.empty((h, ncomp))
            s = np.empty(ncomp)
            genes = np.empty(h, dtype=np.int32)
        check(self.lib.adobo_fetch_pca(self.h, comp, s, contr, genes))
        return dict(comp=comp, contr=contr, s=s, genes=genes if out is not None else genes.astype(np.int64))

    def fetch_knn(self, k=None):
        sz = self.result_sizes()
        if k is not None and k != sz['k']:
            raise Exception('fetch_knn: the device holds k=%d neighbours, not %d' % (sz['k'], k))
        out = np.empty((sz['knn_n'], sz['k']), dtype=np.int64)
        check(self.lib.adobo_fetch_knn(self.h, out))
        return out

    def fetch_edges(self, ne=None, counts=False, out=None):
        """Edge list of the last SNN (E x 2 int32); with counts=True also V_ij per edge.  out: (src, dst) caller-owned
        int32 arrays of at least E entries (returned as they are, no stacking)."""
        have = self.result_sizes()['n_edges']
        if ne is not None and ne != have:
            raise Exception('fetch_edges: the device holds %d edges, not %d' % (have, ne))
        ne = have
        if out is not None:
            src, dst = out
            if src.size < ne or dst.size < ne:
                raise Exception('fetch_edges: out arrays hold fewer than %d entries' % ne)
            check(self.lib.adobo_snn_fetch(self.h, src, dst))
            return src[:ne], dst[:ne]
        if counts:
            v = np.empty(max(ne, 1), dtype=np.int32)
            check(self.lib.adobo_snn_fetch_counts(self.h, v))
        src = np.empty(max(ne, 1), dtype=np.int32)
        dst = np.empty(max(ne, 1), dtype=np.int32)
        check(self.lib.adobo_snn_fetch(self.h, src, dst))
        e = np.stack([src[:ne], dst[:ne]], axis=1)
        return (e, v[:ne]) if counts else e


_default = None


def default_engine():
    """Process-wide engine on ADOBO_B200_DEVICE / LOCAL_RANK / device 0."""
    global _default
    if _default is None:
        _default = Engine()
    return _default


def set_default_engine(e):
    global _default
    _default = e
