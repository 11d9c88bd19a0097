"""Oracle restatement of the tip-data layout steps of /root/reference/vbsky/fasta.py
(test infrastructure): ``_process_tips`` (fasta.py:539-548) and ``pad_tips`` (fasta.py:325-331).
"""
import numpy as np

from .substitution import encode_partials
from .util import TipData


def process_tips(sids, seqs, node_mapping):
    """fasta.py:539-548.  ``sids[i]`` is the (already remapped) name of sequence i, ``seqs[i]`` its
    string, ``node_mapping`` the name -> node index map returned with the TreeData.
    Returns (TipData(partials [S,n,4], counts [S]), S)."""
    tip_partials0 = np.array([encode_partials(seq) for seq in seqs]).transpose([1, 0, 2])
    perm = np.argsort([node_mapping[s] for s in sids])
    tip_partials = tip_partials0[:, perm]
    tip_data_c = TipData(*np.unique(tip_partials, axis=0, return_counts=True))
    return tip_data_c, tip_data_c.partials.shape[0]


def pad_tips(tip_data_cs):
    """fasta.py:325-331: pad every tree's patterns to the ensemble maximum with all-ones rows, count 0."""
    max_partial_count = np.max([t.partials.shape[0] for t in tip_data_cs])
    out = []
    for t in tip_data_cs:
        pad_size = max_partial_count - t.partials.shape[0]
        pad = np.tile([1, 1, 1, 1], (pad_size, t.partials.shape[1], 1))
        partials = np.concatenate([t.partials, pad])
        counts = np.concatenate([t.counts, np.zeros(pad_size)])
        out.append(TipData(partials, counts))
    return out


def read_nexus_matrix(path):
    """Tiny reader for the MATRIX block of a non-interleaved NEXUS file (example/hcv.nexus)."""
    names, seqs = [], []
    in_matrix = False
    with open(path) as fh:
        for line in fh:
            s = line.strip()
            if not in_matrix:
                if s.upper() == "MATRIX":
                    in_matrix = True
                continue
            if s == ";" or s.upper().startswith("END"):
                break
            if not s:
                continue
            parts = s.split()
            names.append(parts[0])
            seqs.append("".join(parts[1:]).rstrip(";"))
    return names, seqs
