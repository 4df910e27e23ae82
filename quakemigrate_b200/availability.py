# -*- coding: utf-8 -*-
"""
``<run>/detect/availability/<year>_<julday>_StationAvailability.csv`` -- the second output of the
detect loop (``_continuous_compute``, quakemigrate/signal/scan.py:431-467), mirrored from
``quakemigrate/io/availability.py``: one row per timestep (index = timestamp of the step as
``UTCDateTime`` prints it, ``YYYY-MM-DDTHH:MM:SS.ffffffZ``), one 0/1 column per ``station_phase``
in the phase-outer / station-inner order of scan.py:433-438, files split by Julian day.

Plain csv, no obspy / pandas needed to write; pinned byte-for-byte by the reference's golden
``examples/benchmarks/Icequake_Iceland/2014_180_StationAvailability.csv``
(``tests/test_scanmseed.py``).
"""

from __future__ import annotations

import csv

from .scanmseed import _julday, _run_path, from_us, to_us

__all__ = ["write_availability", "read_availability", "NoStationAvailabilityDataException"]


class NoStationAvailabilityDataException(Exception):
    """``util.NoStationAvailabilityDataException``."""

    def __init__(self):
        super().__init__("QuakeMigrate Error: No .StationAvailability files found.")


def _stamp(us: int) -> str:
    return from_us(us).strftime("%Y-%m-%dT%H:%M:%S.%f") + "Z"


def write_availability(run, times, rows, columns=None):
    """
    availability.py:123-152.  ``times``: one timestamp per timestep; ``rows``: one dict
    ``{"STATION_PHASE": 0/1}`` per timestep, or ``None`` for a step without data (written as
    zeros, scan.py:459-461); ``columns``: column order (default: the keys of the first dict).
    Returns the list of files written.
    """
    if columns is None:
        first = next((r for r in rows if r is not None), None)
        if first is None:
            raise ValueError("no availability information to take the columns from")
        columns = list(first.keys())
    fpath = _run_path(run) / "detect" / "availability"
    fpath.mkdir(exist_ok=True, parents=True)
    by_day: dict = {}
    for t, row in zip(times, rows):
        us = to_us(t)
        by_day.setdefault(_julday(us), []).append((us, row))
    files = []
    for (year, julday), entries in by_day.items():
        file = fpath / f"{year}_{julday:03d}_StationAvailability.csv"
        with open(file, "w", newline="") as f:
            w = csv.writer(f, lineterminator="\n")
            w.writerow([""] + list(columns))
            for us, row in entries:
                w.writerow([_stamp(us)] + [int(row.get(c, 0)) if row is not None else 0 for c in columns])
        files.append(file)
    return files


def read_availability(run, starttime, endtime):
    """
    availability.py:21-73: ``(index, columns, table)`` -- timestamps (``datetime``), column names
    and an int array ``(n_steps, n_columns)`` -- from the day files between the two timestamps.
    Old-style files (one column per station) are expanded to ``station_P`` / ``station_S``
    (availability.py:76-120).
    """
    import numpy as np

    fpath = _run_path(run) / "detect" / "availability"
    day = 86_400_000_000
    t = (to_us(starttime) // day) * day
    index, table, columns = [], [], None
    while t <= to_us(endtime):
        year, julday = _julday(t)
        file = fpath / f"{year}_{julday:03d}_StationAvailability.csv"
        if file.exists():
            with open(file, newline="") as f:
                rows = list(csv.reader(f))
            cols = rows[0][1:]
            body = [[int(v) for v in r[1:]] for r in rows[1:]]
            if cols and len(cols[0].split("_")) != 2:  # old format
                body = [r + r for r in body]
                cols = [f"{c}_P" for c in cols] + [f"{c}_S" for c in cols]
            columns = columns or cols
            import datetime as _dt
            index += [_dt.datetime.strptime(r[0], "%Y-%m-%dT%H:%M:%S.%fZ") for r in rows[1:]]
            table += body
        t += day
    if columns is None:
        raise NoStationAvailabilityDataException
    return index, columns, np.array(table, dtype=int)
