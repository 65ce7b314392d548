"""adobo_b200 -- B200 (sm_100a) implementation of adobo's cell-embedding hot path behind adobo's
own API:  normalize.norm -> hvg.find_hvg -> dr.pca(method='irlb') -> clustering.generate (knn/snn).

    import adobo_b200 as ad
    exp = ad.data.dataset(counts_frame)          # or a reference adobo.data.dataset object
    ad.normalize.norm(exp); ad.hvg.find_hvg(exp); ad.dr.pca(exp); ad.clustering.generate(exp, clust_alg=None)

All numerics run in libadobo_b200.so (hand-written CUDA, C-ABI in include/adobo_b200.h).  The
package imports without a GPU, but every computing call raises when no CUDA device / library
is available -- there is no CPU fallback.
"""
from . import data, normalize, hvg, dr, irlbpy, clustering, preproc, plotting, engine, synth  # noqa: F401

__version__ = '0.1.0'
