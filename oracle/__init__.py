"""CPU oracle for the VBSKY hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A literal fp64 restatement (NumPy, PyTorch-fp64 for autograd, plus a C/OpenMP
port of the two pruning scans) of the arithmetic in the reference
``/root/reference/vbsky`` for the path named by BASELINE.json:north_star:

    prune.py:15-150, substitution.py:10-92, bdsky.py:18-231,268-345,
    tree_data.py:55-117,256-376,411-517, util.py:13-71, optim.py:19-32,
    fasta.py:325-331,539-548

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import anything from this package.  Nothing under
``vbsky_b200/`` imports it; the product path fails loudly without its CUDA library.

Pinning status (see DESIGN.md "Oracle"):
* tree indexing, ``height_transform``, ``inverse_height_transform``,
  ``lower_sampling_times``, ``order_events``: pinned by the reference's own
  doctest values (tests/test_oracle_goldens.py).
* ``prune_loglik``, its gradient, ``_tree_loglik``, ``loglik``: the reference holds
  no golden vector or test for them and cannot be imported here (JAX, msprime,
  tskit, Bio are absent and there is no network) -> **parity unpinned** by reference
  outputs.  The restatement is instead validated by brute-force enumeration over
  internal states, closed-form JC69 likelihoods, central finite differences and
  torch.autograd (fp64).
* third-party arithmetic restated from its published definition: ``msprime.HKY``
  (setup.py:15 ``msprime>=1.1.0``, unpinned; call site substitution.py:75-78).
"""
