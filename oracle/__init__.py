"""TEST INFRASTRUCTURE ONLY.

`oracle/` holds a CPU restatement of the reference's algorithm for the batched
kinematic motion-query path (see `parc_oracle.py`).  It is the *checker*:

  * only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` /
    `--impl reference` legs of `bench.py` may import it;
  * nothing under `parc_b200/` imports it, and the product path raises if the
    CUDA library is missing rather than falling back to anything here.

The reference is pure Python/PyTorch (no C/C++/CUDA sources), so there is no C
restatement to compile and no `oracle/_ref` binary: the restatement is written
against `torch` CPU tensors (the reference's own arithmetic dependency) and was
pinned against the imported reference in the authoring container by
`oracle/make_golden.py`, whose outputs are committed under `tests/golden/`.
"""
