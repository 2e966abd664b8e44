"""CPU oracle for the StrataNet2 PointNet++ / 2D-projection hot path.

TEST INFRASTRUCTURE -- not product code.  Only tests/, ``__graft_entry__.smoke()`` and
bench.py's ``cpu_baseline`` / ``--impl reference`` legs may import this package.  The product
package (stratanet2-vegetation-coverage-maps_b200/) never imports it and fails loudly when its
CUDA library is missing.

Contents
  thirdparty_ops.py   restated torch_cluster / torch_scatter / torch_geometric ops
  csrc/sn2_oracle.c   plain-C fps / radius / knn used by thirdparty_ops (gcc, -ffp-contract=off)
  pointnet2_port.py   restated PointNet2.forward + both projections (reference file:line cited)
  ref_loader.py       runs the reference's OWN model files verbatim on thirdparty_ops
                      (from /root/reference here, or from the git-ignored staging oracle/_ref/)
  stage_ref.py        stages the two reference files into oracle/_ref/ (never committed)

Parity status: the third-party arithmetic is "parity unpinned" against the real wheels (they are
not installable in this environment); it is pinned by tests/golden/ (reference files run verbatim
here, generator committed) and by known-answer cases.
"""
