"""CPU oracle for the AMUSE gesture-sampling hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``amuse_b200/`` may import this
package: only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` are allowed to, and there only as the
checker (or as the timed CPU baseline), never as the product path.

What is in here
---------------
``weights``          deterministic synthetic state-dicts carrying the reference's
                     own key names / shapes -- the reference
                     ships no checkpoints, so parity is defined on seeded weights.
``lpdm_ref``         plain-PyTorch (CPU, fp32 or fp64) restatement of
                     ``PretrainedLPDM_v1.diffusion_backward`` and everything under
                     it: Denoiser, DDIM / DDPM scheduler steps, MotionPrior.decode,
                     6D -> axis-angle.  Every function cites the reference
                     file:line it follows.
``ast_ref``          restatement of the three AST (DeiT-base) audio encoders and of
                     ``process_single_seq``.
``reference_loader`` imports the reference's OWN ``Denoiser`` / ``MotionPrior`` /
                     rotation code from ``/root/reference`` (build container only;
                     that tree does not exist on the GPU box).
``make_golden``      runs the reference's own modules on seeded inputs and writes
                     the small fixtures under ``tests/golden/``.

Pinning status
--------------
* Denoiser, MotionPrior.decode, 6D->axis-angle: PINNED -- the restatement is checked
  (tests/test_oracle_vs_reference.py, build container) against the reference's own
  modules, and the committed golden vectors were produced by those modules.
* DDIM / DDPM scheduler (diffusers==0.17.1, amuse.yml:143) and the AST encoders
  (timm==0.4.5, amuse.yml:281): PARITY UNPINNED -- neither library is present in the
  build container and the reference repo has no tests or golden vectors for them;
  they are restated from the published algorithms and anchored on the reference's
  call sites (infer_ldm.py:116-125,142-161; audio_main_new.py:174-204).
"""
