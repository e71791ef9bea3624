"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the rbvfit likelihood hot path.

Nothing in ``rbvfit_b200/`` may import this package.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` use it, and only as the checker / the CPU arm being timed.

* ``voigt_oracle``  numpy + ``scipy.special.wofz`` restatement of the reference algorithm
                    (each function cites the reference file:line it follows).
* ``refshim``       imports the UNMODIFIED reference from ``/root/reference/src`` behind a
                    small astropy/emcee/matplotlib shim (build container only; the GPU box
                    has no ``/root/reference``).
* ``make_golden``   runs the real reference through ``refshim`` and writes the committed
                    fixtures under ``tests/golden/``.
* ``stretch_replay`` numpy restatement of the device-resident stretch-move sampler
                    (``rbv_stretch_run``), Philox random streams included.
* ``slice_replay``  numpy restatement of the device-resident ensemble slice sampler
                    (``rbv_slice_run``), same random streams.

Parity status: the reference ships no golden vectors or value-asserting tests for this path
(SURVEY.md section 4), so the oracle is pinned against outputs of the reference code itself,
executed in the build container by ``make_golden.py`` (fixtures committed).  The astropy
kernel construction (``Gaussian1DKernel``, ``convolve(boundary='extend')``) is restated from
astropy's documented behaviour because astropy is not installed here: that one input is
"parity unpinned" and says so in DESIGN.md.
"""
