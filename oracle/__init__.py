"""CPU oracle for the physics-inspired ultrasound augmentation chain.

TEST INFRASTRUCTURE ONLY.  Nothing under ``ultrasoundaugmentation_b200/`` may import this
package; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs use it, and there only as the checker or the timed CPU arm.

PARITY UNPINNED.  The mounted reference (/root/reference) is README.md:1-2 only: a title
and the tagline "Realistic and Physics based Ultrasound Augmentation".  It ships no code,
tests, fixtures or golden vectors, so this oracle is SELF-WRITTEN from the method
definitions in SURVEY.md Appendix A (A.0-A.5).  Every function cites the appendix clause
it restates (there is no reference file:line to cite beyond README.md:2, which only names
the topic).  What pins the oracle instead are closed-form known-answer tests
(tests/test_oracle_kat.py): identity warps, the exact linear-ramp confidence map of a
constant image, boundary rows, direct-solve residuals, and cross-checks against
scipy.ndimage.map_coordinates.

Modules
-------
synth        seeded synthetic B-mode speckle + bone arc + shadow
deform       A.1 bone-surface scan / smoothing / displacement, A.2 remap + TGC
confmap      A.3 random-walk confidence map (direct sparse solve, fp64)
snr_reverb   A.4 SNR change, A.5 reverberation
chain        A.1-A.5 in chain order
"""

from . import synth, deform, confmap, snr_reverb, chain  # noqa: F401
