"""``Data``: a thin pandas-backed container with one property per declared column
(mirrors /root/reference/napkon_string_matching/types/data.py:15-105; host side only)."""
from __future__ import annotations

import json
from hashlib import md5
from pathlib import Path
from typing import Iterable, List

import pandas as pd


def gen_hash(string: str) -> str:
    return md5(string.encode("utf-8"), usedforsecurity=False).hexdigest()


def _column_property(column: str) -> property:
    def fget(self):
        return self._data[column]

    def fset(self, value):
        frame = self._data
        frame[column] = value

    return property(fget, fset)


class Data:
    """Wraps a DataFrame; unknown attributes fall through to it.  Subclasses list their columns
    as an Enum in ``__columns__`` and get ``obj.<name.lower()>`` accessors for each."""

    __slots__ = ["_data"]
    __columns__: Iterable = []
    __column_names__: List[str] = []

    def __init_subclass__(cls, **kwargs):
        super().__init_subclass__(**kwargs)
        cls._install_column_properties()

    @classmethod
    def _install_column_properties(cls):
        columns = list(cls.__columns__)
        for column in columns:
            setattr(cls, column.name.lower(), _column_property(column.value))
        if columns:
            cls.__column_names__ = [column.value for column in columns]

    def __init__(self, data=None):
        self._data = data._data if isinstance(data, Data) else pd.DataFrame(data)

    # ---- DataFrame pass-through -------------------------------------------------------
    def __getattr__(self, name: str):
        if name == "_data":
            raise AttributeError(name)
        return getattr(self._data, name)

    def __getitem__(self, key):
        result = self._data[key]
        return self.__class__(result) if isinstance(result, pd.DataFrame) else result

    def __setitem__(self, key, value):
        self._data[key] = value

    def __len__(self) -> int:
        return len(self._data)

    def __repr__(self) -> str:
        return repr(self._data)

    __str__ = __repr__

    def __eq__(self, other) -> bool:
        return self._data.equals(other._data if isinstance(other, Data) else other)

    def dataframe(self) -> pd.DataFrame:
        return self._data

    def dropna(self, *args, **kwargs):
        return self.__class__(self._data.dropna(*args, **kwargs))

    def drop(self, *args, **kwargs):
        return self.__class__(self._data.drop(*args, **kwargs))

    def merge(self, *args, **kwargs):
        return self.__class__(self._data.merge(*args, **kwargs))

    def drop_superfluous_columns(self, columns: List[str] | None = None) -> None:
        keep = set(columns if columns is not None else self.__column_names__)
        self._data.drop(columns=[c for c in self._data.columns if c not in keep], inplace=True)

    # ---- serialisation ----------------------------------------------------------------
    def to_csv(self) -> str:
        return self._data.to_csv(index=False)

    def to_json(self, *args, **kwargs):
        return self._data.to_json(*args, **kwargs)

    def hash(self) -> str:
        return gen_hash(self.to_csv())

    def get_items(self):
        return [("Sheet1", self._data)]

    @classmethod
    def read_json(cls, file_name: str | Path, *args, **kwargs):
        payload = json.loads(Path(file_name).read_text(encoding="utf-8"))
        result = cls(data=payload)
        result._data.reset_index(drop=True, inplace=True)
        return result

    def write_json(self, file_name: str | Path, *args, **kwargs) -> None:
        Path(file_name).write_text(self.to_json(orient="records", indent=4), encoding="utf-8")

    def write_csv(self, file_name: str | Path, *args, **kwargs) -> None:
        Path(file_name).write_text(self.to_csv(), encoding="utf-8")
