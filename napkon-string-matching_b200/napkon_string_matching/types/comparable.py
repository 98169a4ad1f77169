"""``Comparable``: the result of one comparison — a frame with ``{Left}{Identifier, Argument,
Variable, Sheet}``, ``{Right}...`` and ``MatchScore`` (Q10; mirrors
/root/reference/napkon_string_matching/types/comparable.py:16-162; host side only)."""
from __future__ import annotations

import json
from enum import Enum
from pathlib import Path
from typing import Dict, List

import pandas as pd

from napkon_string_matching.types.data import Data


class Columns(Enum):
    IDENTIFIER = "Identifier"
    PARAMETER = "Parameter"
    VARIABLE = "Variable"
    SHEET = "Sheet"
    MATCH_SCORE = "MatchScore"


QUESTION_OUTPUT = "Argument"
COLUMN_NAMES = [Columns.IDENTIFIER.value, QUESTION_OUTPUT, Columns.VARIABLE.value,
                Columns.SHEET.value]
LEFT_NAME, RIGHT_NAME, DATA_NAME = "left_name", "right_name", "data"
_OWN = ("left_name", "right_name", "data")


class Comparable:
    """``comp.match_<col>`` addresses the left side, ``comp.<col>`` the right side and
    ``comp.match_score`` the score, as in the reference."""

    def __init__(self, data=None, left_name: str | None = None, right_name: str | None = None):
        if left_name is None or right_name is None:
            if not (isinstance(data, dict) and {LEFT_NAME, RIGHT_NAME, DATA_NAME} <= set(data)):
                raise AttributeError(
                    "Either provide 'left_name' AND 'right_name' or a dictionary in 'data' providing "
                    f"the entries {LEFT_NAME}, {RIGHT_NAME} AND {DATA_NAME}")
            left_name, right_name, data = data[LEFT_NAME], data[RIGHT_NAME], data[DATA_NAME]
        object.__setattr__(self, "left_name", left_name)
        object.__setattr__(self, "right_name", right_name)
        object.__setattr__(self, "data", Data(data))

    def _column_for(self, name: str) -> str | None:
        parts = name.split("_")
        tail = parts[-1].title()
        if tail in COLUMN_NAMES:
            return (self.left_name if parts[0] == "match" else self.right_name) + tail
        if name == Columns.MATCH_SCORE.name.lower():
            return Columns.MATCH_SCORE.value
        return None

    def __getattr__(self, name: str):
        if name in _OWN:
            raise AttributeError(name)
        column = self._column_for(name)
        return self.data[column] if column is not None else getattr(self.data, name)

    def __setattr__(self, name: str, value) -> None:
        column = self._column_for(name)
        if column is not None:
            self.data[column] = value
        else:
            setattr(self.data, name, value)

    def __getitem__(self, item):
        result = self.data[item]
        if isinstance(result, Data):
            return Comparable(data=result, left_name=self.left_name, right_name=self.right_name)
        return result

    def __len__(self) -> int:
        return len(self.data)

    def __repr__(self) -> str:
        return repr(self.data)

    __str__ = __repr__

    def __eq__(self, other) -> bool:
        return (isinstance(other, Comparable) and self.left_name == other.left_name
                and self.right_name == other.right_name and self.data == other.data)

    def _rewrap(self, frame) -> "Comparable":
        return Comparable(data=frame, left_name=self.left_name, right_name=self.right_name)

    def dropna(self, *args, **kwargs):
        return self._rewrap(self.data.dataframe().dropna(*args, **kwargs))

    def drop(self, *args, **kwargs):
        return self._rewrap(self.data.dataframe().drop(*args, **kwargs))

    def merge(self, *args, **kwargs):
        return self._rewrap(self.data.dataframe().merge(*args, **kwargs))

    def dataframe(self) -> pd.DataFrame:
        return self.data.dataframe()

    def sort_by_score(self) -> None:
        self.data.dataframe().sort_values(by=Columns.MATCH_SCORE.value, ascending=False, inplace=True)

    def drop_superfluous_columns(self, columns: List[str] | None = None) -> None:
        self.data.drop_superfluous_columns(columns)

    def to_json(self, orient: str | None = None, *args, **kwargs) -> str:
        if orient == "records" and not args and set(kwargs) == {"indent"} and isinstance(kwargs["indent"], int):
            fast = _records_payload_json(self.left_name, self.right_name, self.data.dataframe(),
                                         kwargs["indent"])
            if fast is not None:
                return fast
        payload = {LEFT_NAME: self.left_name, RIGHT_NAME: self.right_name,
                   DATA_NAME: self.data.to_dict(orient=orient)}
        return json.dumps(payload, *args, **kwargs)

    @classmethod
    def read_json(cls, file_name: str | Path, *args, **kwargs) -> "Comparable":
        return cls(data=json.loads(Path(file_name).read_text(encoding="utf-8")))

    def write_json(self, file_name: str | Path, *args, **kwargs) -> None:
        Path(file_name).write_text(self.to_json(orient="records", indent=4), encoding="utf-8")


def _records_payload_json(left_name, right_name, frame: pd.DataFrame, indent: int) -> str | None:
    """The text ``json.dumps({left_name, right_name, data: frame.to_dict("records")}, indent=n)``
    produces (the reference's cache file, comparable.py:61-67 + writable_json.py:20), built column
    by column: every column is encoded by ONE call of the C encoder (an indented ``json.dumps``
    runs the pure-Python encoder: 18 s for 200k kept pairs) and the lines are assembled with
    string joins.  Returns None for frames it does not cover (nested cells, no columns, indent
    <= 0); the caller then takes the generic path."""
    if indent <= 0 or frame.shape[1] == 0 or not all(isinstance(c, str) for c in frame.columns):
        return None
    pad1, pad2, pad3 = " " * indent, " " * (2 * indent), " " * (3 * indent)
    head = ("{\n" + pad1 + json.dumps(LEFT_NAME) + ": " + json.dumps(left_name) + ",\n"
            + pad1 + json.dumps(RIGHT_NAME) + ": " + json.dumps(right_name) + ",\n"
            + pad1 + json.dumps(DATA_NAME) + ": ")
    n = len(frame)
    if n == 0:
        return head + "[]\n}"
    columns = []
    last = frame.shape[1] - 1
    for j, name in enumerate(frame.columns):
        values = frame[name].tolist()
        try:
            # "\x00" cannot occur inside an encoded value (control characters are escaped)
            text = json.dumps(values, separators=("\x00", ": "))[1:-1]
        except (TypeError, ValueError):
            return None
        if text[:1] in "[{" or "\x00[" in text or "\x00{" in text:
            return None  # a list / dict cell: its inner layout depends on the indent
        encoded = text.split("\x00")
        if len(encoded) != n:
            return None
        prefix = (pad2 + "{\n" if j == 0 else "") + pad3 + json.dumps(name) + ": "
        suffix = ",\n" if j < last else "\n" + pad2 + "}"
        columns.append([prefix + e + suffix for e in encoded])
    records = map("".join, zip(*columns))
    return head + "[\n" + ",\n".join(records) + "\n" + pad1 + "]\n}"


class ComparisonResults:
    """name -> Comparable; one sheet per comparison when written."""

    def __init__(self, comp_dict: Dict[str, Comparable] | None = None) -> None:
        self.results = comp_dict if comp_dict else {}

    def __setitem__(self, item, value):
        self.results[item] = value

    def __getitem__(self, item):
        return self.results[item]

    def items(self):
        return self.results.items()

    get_items = items

    def write_excel(self, file: str) -> None:
        """xlsx through pandas/openpyxl like the reference; where openpyxl is missing, one JSON
        document with the same sheets."""
        path = Path(file)
        path.parent.mkdir(parents=True, exist_ok=True)
        try:
            import openpyxl  # noqa: F401
        except ImportError:
            sheets = {name: json.loads(comp.to_json(orient="records")) for name, comp in self.items()}
            path.with_suffix(".json").write_text(json.dumps(sheets, indent=1), encoding="utf-8")
            return
        with pd.ExcelWriter(file, engine="openpyxl") as writer:
            for name, comp in self.items():
                comp.dataframe().to_excel(writer, sheet_name=name, index=False)
