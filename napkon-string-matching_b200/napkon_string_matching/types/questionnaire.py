"""Cohort questionnaire items (``Datensatztabelle`` rows) as comparable data.  Only what the
comparison path needs: column set, ``Term = [*header, question, parameter]``
(/root/reference/napkon_string_matching/types/questionnaire.py:13-68).  The Excel parser of the
reference (types/dataset_table/) is ETL and out of scope; items arrive as JSON / DataFrames."""
from __future__ import annotations

from enum import Enum
from typing import List

import pandas as pd

import napkon_string_matching.types.comparable as comp
from napkon_string_matching.types.comparable_data import ComparableColumns, ComparableData


class Columns(Enum):
    SHEET = "Sheet"
    FILE = "File"
    HEADER = "Header"
    QUESTION = "Question"
    OPTIONS = "Options"
    VARIABLE = "Variable"
    PARAMETER = "Parameter"
    UID = "Uid"
    CATEGORY = "Category"


class Questionnaire(ComparableData):
    __columns__ = list(ComparableColumns) + list(Columns)
    __category_column__ = Columns.CATEGORY.value
    __column_mapping__ = {Columns.PARAMETER.value: comp.Columns.PARAMETER.value}

    def concat(self, others: List["Questionnaire"]):
        if not others:
            return self
        if not isinstance(others[0], Questionnaire):
            raise TypeError("'other' should be of type '{}' but is of type '{}'".format(
                type(self).__name__, type(others[0]).__name__))
        return self.__class__(pd.concat([self._data, *[o._data for o in others]], ignore_index=True))

    def add_terms(self, language: str = "german"):
        self.term = [
            self.gen_term(*(header or []), question, parameter)
            for header, question, parameter in zip(self.header, self.question, self.parameter)
        ]
