# -*- coding: utf-8 -*-
"""
quakemigrate_b200 -- B200-native (sm_100a) implementation of QuakeMigrate's coalescence hot path.

Scope: the ``migrate`` / ``find_max_coa`` / STA-LTA C extension that
``QuakeScan._compute`` calls (``quakemigrate/signal/scan.py:594-647``), behind the reference's
own interfaces:

``quakemigrate_b200.core``    drop-in for ``quakemigrate.core`` (same five functions) + the
                              device-resident :class:`~quakemigrate_b200.core.Engine`
``quakemigrate_b200.onsets``  ``Onset`` / ``OnsetData`` / ``STALTAOnset`` plugin API
``quakemigrate_b200.lut``     ``serve_traveltimes`` / ``index2coord`` contract, reference LUT reader
``quakemigrate_b200.scan``    ``QuakeScan._compute`` (detect) and the marginal-window locate
``quakemigrate_b200.sharding``node-slab partitioning for multi-GPU runs

Importing never touches the GPU; calling anything that computes requires ``libqmb200.so`` and a
B200 -- there is no CPU fallback.
"""

__version__ = "0.1.0"

from . import synthetic  # noqa: F401  (host-only helpers)

__all__ = ["core", "onsets", "lut", "scan", "sharding", "synthetic", "__version__"]
