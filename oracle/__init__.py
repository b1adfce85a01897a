"""CPU oracle for the OT knowledge-distillation hot path -- TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is product code.  Only ``tests/``, ``__graft_entry__.smoke()``
and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it, and
only as the checker / the timed CPU baseline.  The product package
(``kd_6d_pose_adlp_b200``) never imports this package and fails loudly when its CUDA
library is missing.

Parity status: the Sinkhorn arithmetic of the reference lives in the third-party
package ``geomloss==0.2.4`` (reference ``requirements.txt:45``), which is absent from
``/root/reference`` and from this image (no network).  The reference ships no tests or
golden vectors for this path.  ``oracle/geomloss_ref.py`` is therefore a restatement of
geomloss' published algorithm: **parity unpinned** against real geomloss outputs.  It is
pinned only by (i) closed-form known answers, (ii) its fp64 self-consistency
(autograd vs analytic backward) and (iii) the reference's OWN in-tree code
(``losses/loss_libs.py``, ``losses/kd_loss.py``, ``postprocess/postprocess_kd.py``)
driven through it in this container to generate ``tests/golden/`` fixtures.
"""
