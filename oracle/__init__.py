"""CPU oracle of the mvlm hot path -- TEST INFRASTRUCTURE ONLY.

Nothing under mvlm_b200/ may import this package.  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use
it, and only as the checker or the timed CPU baseline.

Parity status: the reference (cvjena/mvlm) ships NO tests, golden vectors or
known-answer fixtures, so the pins are produced by running
the reference's own importable modules in the build container
(tools/make_golden.py -> tests/golden/*.npz):
  * hourglass CNN, peaks, rays, LSQ/RANSAC consensus: pinned against the
    reference code executed verbatim (oracle/ref_loader.py).
  * renderer and surface snap: the reference delegates to VTK, which is not
    installable here -> restated from the reference call sites
    (render3d.py:53-77,114-177; estimator3d.py:252-285); PARITY UNPINNED for
    these two stages (pinned only by closed-form properties).
"""
