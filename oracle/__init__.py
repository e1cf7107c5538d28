"""CPU oracle for the Attentive-VAE attention + ELBO hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``deep_attentive_vi_b200/`` may import
this package; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs do, and there only as the checker
or as the timed CPU arm.

PARITY UNPINNED: the reference (ifiaposto/Deep_Attentive_VI) ships no tests, no
golden vectors and cannot be imported here (TensorFlow 2.3 / TFP 0.11 / TFA 0.13
are not installed and there is no network).  The only known answers the
reference publishes are the trainable-parameter counts in its README
(118.97 M / 7.87 M); ``tests/test_model_structure.py`` pins the host model to
them.  Everything else is an op-for-op restatement of the reference's Python
(each function cites the file:line it follows) plus the library semantics listed
in SURVEY.md section 8c, cross-checked by closed-form invariants.
"""
