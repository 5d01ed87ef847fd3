"""``Data2D`` — the result container of the simulator (boundary subset of the reference class).

``data`` has shape (n_depth, n_time); the npz keys are ``data, taxis, daxis, start_time`` — the
format ``PDS1D_*.pack_result`` writes (/root/reference/src/fiberis/simulator/core/pds.py:401-415)
and ``Data2D.load_npz`` reads (/root/reference/src/fiberis/analyzer/Data2D/core2D.py:197-279).
Filtering / merging / plotting are DAS post-processing and outside the hot path.
"""
from __future__ import annotations

import datetime
import os
from typing import Optional, Tuple, Union

import numpy as np

from fiberis_b200.utils.history_utils import InfoManagementSystem

_TIME_FORMATS = ("%Y-%m-%d %H:%M:%S.%f", "%Y-%m-%d %H:%M:%S")


def _parse_start_time(raw):
    if isinstance(raw, np.ndarray) and raw.size == 1:
        raw = raw.item()
    if isinstance(raw, np.datetime64):
        return raw.astype("datetime64[ms]").astype(datetime.datetime)
    if isinstance(raw, datetime.datetime):
        return raw
    if isinstance(raw, str):
        try:
            return datetime.datetime.fromisoformat(raw)
        except ValueError:
            for fmt in _TIME_FORMATS:
                try:
                    return datetime.datetime.strptime(raw, fmt)
                except ValueError:
                    continue
        raise ValueError(f"Could not parse start_time string: {raw}")
    return None  # unknown type: the reference leaves start_time untouched


class Data2D:
    def __init__(self, data: Optional[np.ndarray] = None, taxis: Optional[np.ndarray] = None,
                 daxis: Optional[np.ndarray] = None, start_time: Optional[datetime.datetime] = None,
                 name: Optional[str] = None):
        self.data, self.taxis, self.daxis = data, taxis, daxis
        self.start_time, self.name = start_time, name
        self.history = InfoManagementSystem()
        self.history.add_record(f"Initialized Data2D object with name: {name}" if name
                                else "Initialized empty Data2D object.")
        if data is not None:  # a mismatch is only logged at construction time
            if taxis is not None and data.shape[1] != taxis.shape[0]:
                self.history.add_record(f"Initialization Warning: Data shape[1] ({data.shape[1]}) does not match "
                                        f"taxis length ({taxis.shape[0]}).", level="WARNING")
            if daxis is not None and data.shape[0] != daxis.shape[0]:
                self.history.add_record(f"Initialization Warning: Data shape[0] ({data.shape[0]}) does not match "
                                        f"daxis length ({daxis.shape[0]}).", level="WARNING")

    # -- setters (validated) ----------------------------------------------------------------------
    def set_data(self, data: np.ndarray) -> None:
        if not isinstance(data, np.ndarray) or data.ndim != 2:
            raise TypeError("Data must be a 2D NumPy array.")
        if self.taxis is not None and data.shape[1] != self.taxis.shape[0]:
            raise ValueError(f"New data's time dimension ({data.shape[1]}) does not match existing taxis length "
                             f"({self.taxis.shape[0]}).")
        if self.daxis is not None and data.shape[0] != self.daxis.shape[0]:
            raise ValueError(f"New data's depth dimension ({data.shape[0]}) does not match existing daxis length "
                             f"({self.daxis.shape[0]}).")
        self.data = data
        self.history.add_record("Data attribute updated.")

    def _set_axis(self, which: str, axis: np.ndarray, dim: int) -> None:
        if not isinstance(axis, np.ndarray) or axis.ndim != 1:
            raise TypeError(f"{which} must be a 1D NumPy array.")
        if self.data is not None and self.data.shape[dim] != axis.shape[0]:
            kind = "time" if dim == 1 else "depth"
            raise ValueError(f"New {which} length ({axis.shape[0]}) does not match existing data's {kind} "
                             f"dimension ({self.data.shape[dim]}).")
        setattr(self, which, axis)
        self.history.add_record(f"{which} updated.")

    def set_taxis(self, taxis: np.ndarray) -> None:
        self._set_axis("taxis", taxis, 1)

    def set_daxis(self, daxis: np.ndarray) -> None:
        self._set_axis("daxis", daxis, 0)

    def set_start_time(self, start_time: datetime.datetime) -> None:
        if not isinstance(start_time, datetime.datetime):
            raise TypeError("start_time must be a datetime.datetime object.")
        self.start_time = start_time
        self.history.add_record(f"Start time set to {start_time}.")

    def set_name(self, name: str) -> None:
        self.name = name
        self.history.add_record(f"Name set to '{name}'.")

    # -- I/O ------------------------------------------------------------------------------------
    def load_npz(self, filename: str) -> None:
        path = filename if filename.endswith(".npz") else filename + ".npz"
        z = np.load(path, allow_pickle=True)
        try:
            data, taxis, daxis, raw_start = z["data"], z["taxis"], z["daxis"], z["start_time"]
        except KeyError as exc:
            raise KeyError(f"NPZ file {path} is missing key: {exc}. Expected 'data', 'taxis', 'daxis', "
                           f"'start_time'.") from exc
        for label, arr, nd in (("data", data, 2), ("taxis", taxis, 1), ("daxis", daxis, 1)):
            if arr.ndim != nd:
                raise ValueError(f"Loaded '{label}' must be {nd}D, got {arr.ndim}D.")
        if data.shape[1] != taxis.shape[0]:
            raise ValueError(f"Loaded data time dimension ({data.shape[1]}) mismatch with taxis length "
                             f"({taxis.shape[0]}).")
        if data.shape[0] != daxis.shape[0]:
            raise ValueError(f"Loaded data depth dimension ({data.shape[0]}) mismatch with daxis length "
                             f"({daxis.shape[0]}).")
        self.data, self.taxis, self.daxis = data.astype(float), taxis.astype(float), daxis.astype(float)
        start = _parse_start_time(raw_start)
        if start is not None:
            self.start_time = start
        self.set_name(os.path.basename(path))
        self.history.add_record(f"Successfully loaded data from {path}.")

    def savez(self, filename: str) -> None:
        if self.data is None or self.taxis is None or self.daxis is None:
            raise ValueError("Data, taxis, and daxis must be set before saving.")
        if not filename.endswith(".npz"):
            filename += ".npz"
        np.savez(filename, data=self.data, taxis=self.taxis, daxis=self.daxis, start_time=self.start_time)
        self.history.add_record(f"Data successfully saved to {filename}.")

    # -- getters --------------------------------------------------------------------------------
    def get_value_by_depth(self, depth: float, interpolate: bool = False) -> Optional[np.ndarray]:
        """Time series of the channel nearest to ``depth`` (None when out of range)."""
        if self.daxis is None or self.data is None:
            raise ValueError("daxis and data must be set.")
        if self.daxis.size == 0 or depth < np.min(self.daxis) or depth > np.max(self.daxis):
            return None
        return self.data[int(np.argmin(np.abs(self.daxis - depth))), :]

    def get_value_by_time(self, time_point: Union[datetime.datetime, float, int]) -> Tuple[np.ndarray, float]:
        """(depth trace at the sample nearest to ``time_point``, that sample's time in seconds)."""
        if self.taxis is None or self.data is None:
            raise ValueError("taxis and data must be set.")
        if self.taxis.size == 0:
            raise ValueError("Cannot get value by time, taxis is empty.")
        if isinstance(time_point, (int, float)):
            seconds = float(time_point)
        elif isinstance(time_point, datetime.datetime):
            if self.start_time is None:
                raise ValueError("start_time must be set to use a datetime object as input.")
            seconds = (time_point - self.start_time).total_seconds()
        else:
            raise TypeError(f"Unsupported type for time_point: {type(time_point)}")
        k = int(np.argmin(np.abs(self.taxis - seconds)))
        return self.data[:, k], float(self.taxis[k])

    def get_max_daxis(self) -> Optional[float]:
        return None if self.daxis is None or self.daxis.size == 0 else float(self.daxis[-1])

    def __str__(self) -> str:
        shape = None if self.data is None else self.data.shape
        return f"Data2D(name={self.name!r}, data={shape}, start_time={self.start_time})"
