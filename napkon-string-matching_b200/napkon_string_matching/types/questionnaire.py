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
    """Columns of a cohort's data set table besides the comparable ones."""
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
    """Items of one cohort questionnaire."""

    __columns__ = [*ComparableColumns, *Columns]
    __category_column__ = Columns.CATEGORY.value
    __column_mapping__ = {Columns.PARAMETER.value: comp.Columns.PARAMETER.value}

    def concat(self, others: List["Questionnaire"]):
        """A new questionnaire holding the rows of this one followed by the others'."""
        if not others:
            return self
        wrong = next((o for o in others if not isinstance(o, Questionnaire)), None)
        if wrong is not None:
            raise TypeError(f"'other' should be of type '{type(self).__name__}' but is of type "
                            f"'{type(wrong).__name__}'")
        frames = [self._data] + [o._data for o in others]
        return type(self)(pd.concat(frames, ignore_index=True))

    def add_terms(self, language: str = "german"):
        """``Term = [*header parts, question, parameter]`` with empty parts dropped."""
        terms = []
        for header, question, parameter in zip(self.header, self.question, self.parameter):
            terms.append(self.gen_term(*(header or ()), question, parameter))
        self.term = terms
