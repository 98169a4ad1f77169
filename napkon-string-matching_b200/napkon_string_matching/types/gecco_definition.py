"""GECCO83 / GECCOplus definition rows as comparable data: ``Term = [category, parameter,
choice]`` and ``Variable := Identifier`` when compared
(/root/reference/napkon_string_matching/types/gecco_definition.py:14-61).  Reading the Excel
definitions (gecco_definition_types/) is ETL and out of scope."""
from __future__ import annotations

from enum import Enum

import pandas as pd

import napkon_string_matching.types.comparable as comp
from napkon_string_matching.types.comparable_data import ComparableColumns, ComparableData


class Columns(Enum):
    CATEGORY = "Category"
    PARAMETER = "Parameter"
    CHOICES = "Choices"


class GeccoDefinition(ComparableData):
    __columns__ = list(ComparableColumns) + list(Columns)
    __category_column__ = Columns.CATEGORY.value
    __column_mapping__ = {}

    def map_for_comparable(self) -> pd.DataFrame:
        frame = super().map_for_comparable().copy()
        frame[comp.Columns.VARIABLE.value] = frame[comp.Columns.IDENTIFIER.value]
        return frame

    def concat(self, other: "GeccoDefinition"):
        if not isinstance(other, GeccoDefinition):
            raise TypeError("'other' should be of type '{}' but is of type '{}'".format(
                type(self).__name__, type(other).__name__))
        return GeccoDefinition(pd.concat([self._data, other._data], ignore_index=True))

    def add_terms(self, language: str = "german"):
        self.term = [self.gen_term(category, parameter, choice)
                     for category, parameter, choice in zip(self.category, self.parameter, self.choices)]


class KdsDefinition(GeccoDefinition):
    """MII core data set rows: ``Term = [category, parameter]``
    (/root/reference/napkon_string_matching/types/kds_definition.py:30-68).  The reference loads
    KDS but never compares it (matcher.py:108-122); the GPU path accepts it like any other side."""

    def add_terms(self, language: str = "german"):
        self.term = [self.gen_term(category, parameter)
                     for category, parameter in zip(self.category, self.parameter)]
