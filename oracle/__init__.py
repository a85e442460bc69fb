"""TEST INFRASTRUCTURE ONLY — CPU oracle for the ViewFormer hot path.

Nothing under ``oracle/`` is product code.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import it, and there only as the checker (or as the timed CPU arm), never
as part of the CUDA path that is shipped or measured.

Contents
--------
``ref_loader.py``    imports the *real* reference torch VQGAN from ``/root/reference``
                     (container only; the GPU box has no ``/root/reference``).
``vqgan_oracle.py``  functional torch-CPU restatement of ``viewformer/models/vqgan_th.py``
                     + ``utils_th.py`` (pinned against the real reference by
                     ``tests/test_oracle_vs_reference.py`` and by the golden fixtures).
``migt_oracle.py``   torch-CPU restatement of ``viewformer/models/migt.py`` +
                     ``branching_attention.py`` — pinned to the reference's own sources
                     executed over ``tf_shim.py`` (TensorFlow itself is not installable here).
``tf_shim.py``       torch-backed stand-in for the TensorFlow ops the reference's transformer
                     files call, so that those files run unmodified (container only).
``vq_lookup.c``      plain-C restatement of the codebook nearest-neighbour search
                     (integer index output), built by ``oracle/Makefile``.
``synth.py``         deterministic synthetic weights / inputs shared by tests & bench.
``make_golden.py``   regenerates ``tests/golden/*.npz`` from the real reference.
"""
