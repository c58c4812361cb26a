"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the synthetic-PTA residual hot path.

Nothing in the product package (``pta_replicator_b200``) may import this
package.  Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline``
/ ``--impl reference`` legs of ``bench.py`` use it, and only as the checker /
the CPU arm being timed.

Contents
--------
``refnumpy``   numpy restatement of the reference algorithms (travels to the
               GPU box; every function cites the reference file:line it follows).
``philox``     numpy restatement of the Philox4x32-10 + Box-Muller stream the
               CUDA kernels use in throughput mode.
``refstubs``   ``sys.modules`` stubs (astropy/pint/enterprise/ephem/holodeck)
               and a duck-typed pulsar that let the UNMODIFIED reference under
               ``/root/reference`` run in the authoring container.  Used by
               ``make_golden.py`` to pin ``refnumpy`` and to generate
               ``tests/golden/*.npz``.  ``/root/reference`` does not exist on
               the GPU box, so nothing there imports ``refstubs``.

Parity pinning: ``refnumpy`` is pinned against (a) the reference's own golden
vector ``tests/libstempo_test_residuals_efac_ecorr_rn_gwb_cgw.npz`` (copied as
``tests/golden/libstempo_golden.npz``) and (b) outputs of the unmodified
reference functions run under ``refstubs`` (``tests/golden/ref_*.npz``).
"""
