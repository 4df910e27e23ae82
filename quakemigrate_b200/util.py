# -*- coding: utf-8 -*-
"""
Pad -> sample arithmetic of the reference (``quakemigrate/util.py``), host logic used by
``QuakeScan._compute`` (scan.py:632-633) and ``Onset.pad`` (base.py:64-93).
"""

from __future__ import annotations

import numpy as np


def time2sample(time: float, sampling_rate: int) -> int:
    """``quakemigrate.util.time2sample`` (util.py:152-172)."""
    return int(round(time * int(sampling_rate)))


def trim2sample(time: float, sampling_rate: int) -> float:
    """``quakemigrate.util.trim2sample`` (util.py:248-269)."""
    return int(np.ceil(time * sampling_rate) / sampling_rate * 1000) / 1000
