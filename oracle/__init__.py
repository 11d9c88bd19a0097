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

Pinning status (see DESIGN.md "Oracle"): **pinned to the reference's own source, executed here.**
* ``oracle/refshim`` imports the UNMODIFIED files under /root/reference/vbsky on a torch-fp64 stand-in for the JAX
  subset they use (JAX, msprime, tskit, Bio are not installable offline) and ``tests/test_reference_pin.py`` asserts
  oracle == reference-code to ~1e-13 for ``expm_action`` (both directions), ``HKY``/``JC69``, ``prune_loglik`` per site,
  the custom-JVP gradient, both partial arrays, ``_tree_loglik`` and ``loglik`` (values and gradients w.r.t. every
  argument, every term switch, both grid modes), ``TreeData`` numbering / ``height_transform`` / inverse,
  ``order_events``, ``_process_tips`` / ``pad_tips`` and ``optim.loss`` under ``value_and_grad``.
* the fixtures under tests/golden/ are outputs of those same reference calls (tests/golden/make_golden.py),
  including two iterations of the reference's own ``SeqData.loop`` on the HCV example; they carry the pin to the
  GPU box, where /root/reference does not exist.
* what is NOT the reference underneath its source: the JAX primitives themselves (torch fp64 ops with JAX's
  documented semantics -- clamped gathers, stable sorts, JAX's linspace formula), reverse-mode differentiation
  (torch.autograd of the same primitive ops; the custom JVP is the reference's own rule, transposed), JAX's PRNG
  stream (draws are recorded and replayed), and third-party arithmetic restated from its published definition:
  ``msprime.HKY`` (setup.py:15 ``msprime>=1.1.0``, unpinned; call site substitution.py:75-78),
  ``jax.example_libraries.optimizers.adagrad``, ``Bio.Phylo.read/to_networkx`` (newick topology only).
* additionally pinned by the reference's doctest values (tests/test_oracle_goldens.py) and validated by brute-force
  enumeration, closed-form JC69 likelihoods, central finite differences (tests/test_oracle_selfcheck.py).
"""
