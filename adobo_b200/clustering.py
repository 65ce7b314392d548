"""adobo.clustering drop-in for the graph half: knn(), snn() and generate()'s knn/snn step
(adobo/clustering.py:25-128, 270-363).  Community detection (Leiden / Louvain / igraph) stays on
the host exactly as in the reference and needs the reference's optional packages."""
from collections import Counter

import numpy as np
import pandas as pd

from . import engine as _engine


# clustering.py:47-48 hands `distance` to sklearn's ball tree; these are the ball-tree metrics that need no extra
# parameter (seuclidean / mahalanobis need V / VI, which the reference cannot pass either; 'minkowski' runs with the
# ball tree's default p = 2).  Codes: include/adobo_b200.h ADOBO_METRIC_*.
METRICS = {'euclidean': 0, 'l2': 0, 'minkowski': 0, 'manhattan': 1, 'cityblock': 1, 'l1': 1, 'chebyshev': 2, 'infinity': 2,
           'hamming': 3, 'canberra': 4, 'braycurtis': 5}


def knn(comp, k=10, distance='euclidean', eng=None):
    """clustering.knn (clustering.py:25-52): exact k nearest neighbours of every cell in PCA
    space; returns an (n, k) int64 array of 1-based indices, column 0 being the cell itself.
    `distance`: the parameter-free ball-tree metrics in METRICS (Euclidean on the tensor cores, the
    others by exact fp64 brute force)."""
    if distance not in METRICS:
        raise ValueError("Unrecognized metric '%s'; adobo_b200.clustering.knn supports %s" % (distance, ', '.join(METRICS)))
    eng = eng or _engine.default_engine()
    x = np.ascontiguousarray(np.asarray(comp), dtype=np.float64)
    return eng.knn(x, k, metric=METRICS[distance])


def snn(nn_idx, k=10, prune_snn=0.067, verbose=False, eng=None):
    """clustering.snn (clustering.py:55-128): shared-nearest-neighbour graph, Jaccard-pruned.
    Returns a DataFrame with 0-based `source_node`, `target_node` (both directions and self
    loops, unweighted, like the reference)."""
    eng = eng or _engine.default_engine()
    edges, pruned, total, ne = eng.snn(np.asarray(nn_idx), k, prune_snn)
    perc_pruned = (pruned / total) * 100 if total else 0.0
    if verbose:
        print('%.2f%% (n=%s) of links pruned' % (perc_pruned, '{:,}'.format(pruned)))
    if verbose and perc_pruned > 80:
        print('WARNING: More than 80% of the edges were pruned')
    return pd.DataFrame({'source_node': edges[:, 0].astype(np.int64), 'target_node': edges[:, 1].astype(np.int64)})


def edge_arrays(snn_graph):
    """What the partitioners need from the SNN graph, as arrays: the number of vertices exactly as
    the reference counts them -- distinct values of the first column (clustering.py:155, 196) -- and a C-contiguous
    (E, 2) int32 edge array, which `igraph.Graph.add_edges` takes directly.  The reference builds a Python list of
    tuples with `itertuples` (clustering.py:159-162): minutes at a million cells."""
    cols = snn_graph.columns
    src = np.asarray(snn_graph[cols[0]])
    edges = np.ascontiguousarray(np.stack([src, np.asarray(snn_graph[cols[1]])], axis=1), dtype=np.int32)
    return int(np.unique(src).shape[0]), edges


def _igraph_from_edges(snn_graph):
    import igraph as ig
    n, edges = edge_arrays(snn_graph)
    g = ig.Graph()
    g.add_vertices(n)
    g.vs['name'] = list(range(1, n + 1))                 # clustering.py:158
    g.add_edges(edges)                                   # clustering.py:159-162 without the Python loop
    return g


def leiden(snn_graph, res=0.8, seed=42):
    """clustering.leiden (clustering.py:131-169) -- host, needs leidenalg + igraph."""
    import leidenalg as la
    g = _igraph_from_edges(snn_graph)
    cl = la.find_partition(g, la.RBERVertexPartition, n_iterations=10, resolution_parameter=res, seed=seed)
    return cl.membership


def louvain(snn_graph, res=0.8, seed=42):
    """clustering.louvain (clustering.py:229-267) -- host, needs python-louvain + networkx."""
    import networkx as nx
    import community as community_louvain
    g = _igraph_from_edges(snn_graph)
    partition = community_louvain.best_partition(nx.Graph(g.get_edgelist()), resolution=res, random_state=seed)
    return [partition[i] for i in sorted(partition)]


def generate(obj, k=10, name=None, distance='euclidean', graph='snn', clust_alg='leiden', prune_snn=0.067,
             res=0.8, save_graph=True, seed=42, verbose=False, eng=None):
    """clustering.generate (clustering.py:270-363).  knn + snn run on the GPU; the graph is
    stored in obj.norm_data[name]['graph'] BEFORE the host partitioner is called, so a missing
    leidenalg / python-louvain installation leaves the graph usable.  clust_alg=None skips the
    partitioning (extension)."""
    m = ('leiden', 'louvain', 'walktrap', 'spinglass', 'multilevel', 'infomap', 'label_prop', 'leading_eigenvector')
    if clust_alg is not None and clust_alg not in m:
        raise Exception('Supported community detection algorithms are: %s' % ', '.join(m))
    if not obj.norm_data:
        raise Exception('Run normalization first before running umap. See here: \
https://oscar-franzen.github.io/adobo/adobo.html#adobo.normalize.norm')
    targets = {}
    if name is None or name == '':
        targets = obj.norm_data
    else:
        targets[name] = obj.norm_data[name]
    for l in targets:
        item = targets[l]
        if verbose:
            print('Running clustering on the %s normalization' % l)
        try:
            comp = item['dr']['pca']['comp']
        except KeyError:
            raise Exception('Compute principal components first using adobo.dr.pca(...)')
        nn_idx = knn(comp, k, distance, eng=eng)
        snn_graph = snn(nn_idx, k, prune_snn, verbose, eng=eng)
        if save_graph:
            obj.norm_data[l]['graph'] = snn_graph
        if clust_alg is None:
            obj.set_assay('clustering')
            continue
        if clust_alg == 'leiden':
            cl = leiden(snn_graph, res, seed)
        elif clust_alg == 'louvain':
            cl = louvain(snn_graph, res, seed)
        else:
            raise Exception('adobo_b200: igraph community detection "%s" stays in adobo.clustering.igraph' % clust_alg)
        mem = pd.Series(cl, index=comp.index)
        obj.norm_data[l]['clusters'][clust_alg] = {'membership': mem}
        obj.set_assay('clustering')
        if verbose:
            cd = dict(Counter(cl))
            print('cluster cells')
            for i in sorted(cd):
                print(i, cd[i])
