"""Freezes adobo.IO.load_matrix_market (IO.py:204-266, UNMODIFIED reference, build container only) on a small
10x-style archive: writes tests/golden/mm_small.tar.gz and tests/golden/mm_small_expected.npz.

    python tests/golden/make_io_golden.py
"""
import gzip
import importlib
import io
import os
import sys
import tarfile

import numpy as np
import scipy.sparse as sp
from scipy.io import mmwrite

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_loader  # noqa: E402
from adobo_b200.synth import synth_counts  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def gz_bytes(raw):
    b = io.BytesIO()
    with gzip.GzipFile(fileobj=b, mode='wb', mtime=0) as f:
        f.write(raw)
    return b.getvalue()


def main():
    counts = synth_counts(37, 53, 0.2, 21)                       # cells x genes
    genes_m = sp.coo_matrix(counts.T)
    b = io.BytesIO()
    mmwrite(b, genes_m, field='integer')
    # label files in the two layouts the reference accepts: "id<TAB>symbol" (with quotes) and a bare name
    genes = ['ENSG%05d\t"GENE%d"' % (i, i) for i in range(53)]
    cells = ['CELL%d-1' % i for i in range(37)]
    members = {'matrix.mtx.gz': gz_bytes(b.getvalue()),
               'genes.tsv.gz': gz_bytes(('\n'.join(genes) + '\n').encode()),
               'barcodes.tsv.gz': gz_bytes(('\n'.join(cells) + '\n').encode())}
    path = os.path.join(HERE, 'mm_small.tar.gz')
    with tarfile.open(path, 'w:gz') as tar:
        for name, raw in members.items():
            ti = tarfile.TarInfo(name)
            ti.size = len(raw)
            tar.addfile(ti, io.BytesIO(raw))
    ref_loader.load()
    ref_io = importlib.import_module('adobo.IO')
    df = ref_io.load_matrix_market(path)                         # dense genes x cells frame
    np.savez_compressed(os.path.join(HERE, 'mm_small_expected.npz'), values=np.asarray(df.values),
                        genes=np.array(df.index, dtype=object).astype(str), cells=np.array(df.columns, dtype=object).astype(str))
    print('reference frame', df.shape, 'sum', int(np.asarray(df.values).sum()))


if __name__ == '__main__':
    main()
