"""White-list / black-list "match group" store: ``{id: {group: [identifiers]}}``
(mirrors the parts of /root/reference/napkon_string_matching/types/mapping.py:12-299 the
comparison path uses: filter_by_group :173-176, get_filtered :200-203,
get_all_mapping_for_groups :281-289; host side only)."""
from __future__ import annotations

import json
from pathlib import Path
from typing import Dict, Iterator, List, Tuple


class MappingEntry:
    """Identifiers of several cohorts / definitions that denote the same concept."""

    def __init__(self, data: Dict[str, List[str]] | None = None) -> None:
        self._groups: Dict[str, List[str]] = data if data is not None else {}

    def __getitem__(self, group: str) -> List[str]:
        return self._groups[group]  # KeyError when the entry has no such group (Q7)

    def __setitem__(self, group: str, identifiers: List[str]) -> None:
        self._groups[group] = identifiers

    def get(self, group: str, default=None):
        return self._groups.get(group, default)

    def add(self, group: str, identifier: str) -> None:
        self._groups.setdefault(group, []).append(identifier)

    def update(self, other: "MappingEntry") -> None:
        for group, identifiers in other.dict().items():
            for identifier in identifiers:
                self.add(group, identifier)

    def has(self, group: str, identifier: str) -> bool:
        return identifier in self._groups.get(group, ())

    def dict(self) -> Dict[str, List[str]]:
        return self._groups

    def get_group_names(self) -> List[str]:
        return list(self._groups)

    def get_group_combination(self, group_left: str, group_right: str):
        if group_left in self._groups and group_right in self._groups:
            return self._groups[group_left], self._groups[group_right]
        return None


class Mapping:
    def __init__(self, data: Dict[str, Dict[str, List[str]]] | None = None) -> None:
        self._entries: Dict[str, MappingEntry] = {
            key: MappingEntry(data=dict(groups)) for key, groups in (data or {}).items()}

    def __len__(self) -> int:
        return len(self._entries)

    def __str__(self) -> str:
        """Content-based and stable across processes.  ``ComparableData.compare`` hashes
        ``str(mapping)`` into its cache key; the reference's Mapping has no ``__str__``, so its
        key embeds ``object.__repr__`` — a memory address — and its compare cache can never hit in
        another process.  Same content -> same key here."""
        return "Mapping" + json.dumps(self.dict(), sort_keys=True, ensure_ascii=False)

    def __iter__(self) -> Iterator[Tuple[str, MappingEntry]]:
        return iter(self._entries.items())

    def items(self):
        return self._entries.items()

    def values(self):
        return self._entries.values()

    def get_group(self, id: str) -> MappingEntry | None:
        return self._entries.get(id)

    def set_group(self, id: str, value: MappingEntry) -> None:
        self._entries[id] = value

    def get_group_names(self) -> List[str]:
        return sorted({g for entry in self._entries.values() for g in entry.get_group_names()})

    def filter_by_group(self, group_name: str) -> Dict[str, List[str]]:
        """Entries with a non-empty ``group_name``; KeyError if any entry lacks the group."""
        return {key: entry[group_name] for key, entry in self._entries.items() if entry[group_name]}

    def get_filtered(self, ids: List[str]) -> "Mapping":
        wanted = set(ids)
        result = Mapping()
        result._entries = {key: entry for key, entry in self._entries.items() if key in wanted}
        return result

    def get_all_mapping_for_groups(self, group_left: str, group_right: str
                                   ) -> List[Tuple[List[str], List[str]]]:
        combos = (entry.get_group_combination(group_left, group_right)
                  for entry in self._entries.values())
        return [c for c in combos if c is not None]

    def update(self, other: "Mapping") -> None:
        for key, entry in other.items():
            if key in self._entries:
                self._entries[key].update(entry)
            else:
                self._entries[key] = entry

    def dict(self) -> Dict[str, Dict[str, List[str]]]:
        return {key: entry.dict() for key, entry in self._entries.items()}

    def to_json(self, indent: int | None = None, *args, **kwargs) -> str:
        return json.dumps(self.dict(), indent=indent)

    @classmethod
    def read_json(cls, file_name: str | Path, *args, **kwargs) -> "Mapping":
        return cls(data=json.loads(Path(file_name).read_text(encoding="utf-8")))

    def write_json(self, file_name: str | Path, *args, **kwargs) -> None:
        Path(file_name).write_text(self.to_json(indent=4), encoding="utf-8")
