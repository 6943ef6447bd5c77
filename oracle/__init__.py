"""CPU oracle for the BEVRender deformable-attention hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``bevrender_b200/`` may import this
package; only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` do, and there only as the checker or
the CPU arm that is timed *beside* the CUDA path.

Every function restates, in plain PyTorch-on-CPU / numpy, the algorithm of one
piece of rpl-cmu/bevrender (citations are ``file:line`` relative to the
reference checkout).  The arithmetic of the reference lives in third-party
PyTorch (``F.grid_sample``, ``einsum``, ``softmax``; no version is pinned by
the reference, the installed torch 2.11.0 is the oracle), so the restatement
calls the same library primitives the reference calls and ``oracle/bilinear.py``
restates ``grid_sample`` itself in numpy (published semantics: bilinear,
zero padding, ``align_corners``).

Pinning: the reference ships no tests, golden vectors or fixtures for this path
 -> the oracle is pinned against outputs of the *reference
modules themselves* executed in the authoring container on seeded inputs
(``tests/golden/make_golden.py`` -> ``tests/golden/*.npz``).  Pieces for which
the reference has no implementation at all (multi-level MSDA with L>1/P>1, the
BEV<->map correlation) are marked "parity unpinned" in their module header.
"""
