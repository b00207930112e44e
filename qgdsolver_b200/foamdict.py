"""OpenFOAM dictionary files (controlDict, fvSchemes, fvSolution, thermophysicalProperties, gravitationalProperties ...):
a small reader for the subset the QGD/QHD solvers consume (SURVEY.md 5.6 lists every key).  Host-side harness code.

Grammar handled: `key value ... ;`, nested `key { ... }`, quoted keys (`"(U|e)"`, matched as regular expressions like
OpenFOAM does), lists `( ... )` with an optional size prefix, dimension sets `[0 1 -2 0 0 0 0]`, `$name` macros that refer
to an entry of the same or an enclosing dictionary, C / C++ comments.  `#include`, `#calc`, `#codeStream` are rejected.
"""
from __future__ import annotations

import re
from typing import Any, Dict, Iterator, List, Optional

_TOKEN = re.compile(r'"(?:[^"\\]|\\.)*"|[{}()\[\];]|[^\s{}()\[\];"]+')
_TRUE = {"true", "on", "yes", "y", "t"}
_FALSE = {"false", "off", "no", "n", "f", "none"}


class FoamDictError(ValueError):
    pass


class FoamDict(dict):
    """dict of entries; a value is a FoamDict (sub-dictionary) or a list of tokens (words, numbers as str, nested lists)."""

    def __init__(self, name: str = "", parent: Optional["FoamDict"] = None):
        super().__init__()
        self.name, self.parent = name, parent

    # ---- lookup with OpenFOAM semantics: literal key first, then quoted keys as regular expressions (last one wins)
    def _find(self, key: str):
        if key in self:
            return super().__getitem__(key)
        for k in reversed(list(self.keys())):
            if k.startswith('"') and re.fullmatch(k[1:-1], key):
                return super().__getitem__(k)
        return None

    def found(self, key: str) -> bool:
        return self._find(key) is not None

    def lookup(self, key: str):
        v = self._find(key)
        if v is None:           # dictionary::lookup: FatalIOError "keyword ... is undefined in dictionary ..."
            raise FoamDictError(f'keyword {key} is undefined in dictionary "{self.path()}"')
        return v

    def path(self) -> str:
        return (self.parent.path() + "/" if self.parent is not None and self.parent.name else "") + self.name

    def sub_dict(self, key: str) -> "FoamDict":
        v = self.lookup(key)
        if not isinstance(v, FoamDict):
            raise FoamDictError(f'keyword {key} in "{self.path()}" is not a dictionary')
        return v

    def sub_or_self(self, key: str) -> "FoamDict":
        """QGDCoeffs.C:81-116: the optional `<type>Dict`, else the dictionary itself"""
        v = self._find(key)
        return v if isinstance(v, FoamDict) else self

    def _tokens(self, key: str) -> List[Any]:
        v = self.lookup(key)
        if isinstance(v, FoamDict):
            raise FoamDictError(f'keyword {key} in "{self.path()}" is a dictionary, not a value')
        return v

    def word(self, key: str, default: Optional[str] = None) -> str:
        if default is not None and not self.found(key):
            return default
        t = self._tokens(key)
        if len(t) != 1 or not isinstance(t[0], str):
            raise FoamDictError(f'keyword {key} in "{self.path()}": expected one word, got {t}')
        return t[0].strip('"')

    def scalar(self, key: str, default: Optional[float] = None) -> float:
        if default is not None and not self.found(key):
            return default
        t = [x for x in self._tokens(key) if not (isinstance(x, list) and x and x[0] == "[")]     # skip a dimension set
        if len(t) == 2 and isinstance(t[0], str) and not _is_number(t[0]):                         # dimensioned: name value
            t = t[1:]
        if len(t) != 1 or not isinstance(t[0], str) or not _is_number(t[0]):
            raise FoamDictError(f'keyword {key} in "{self.path()}": expected a scalar, got {t}')
        return float(t[0])

    def label(self, key: str, default: Optional[int] = None) -> int:
        if default is not None and not self.found(key):
            return default
        v = self.scalar(key)
        if v != int(v):
            raise FoamDictError(f'keyword {key} in "{self.path()}": expected an integer, got {v}')
        return int(v)

    def switch(self, key: str, default: Optional[bool] = None) -> bool:
        if default is not None and not self.found(key):
            return default
        w = self.word(key).lower()
        if w in _TRUE:
            return True
        if w in _FALSE:
            return False
        raise FoamDictError(f'keyword {key} in "{self.path()}": {w} is not a Switch')

    def vector(self, key: str) -> List[float]:
        t = [x for x in self._tokens(key) if not (isinstance(x, list) and x and x[0] == "[")]
        t = [x for x in t if isinstance(x, list)] or t
        v = t[-1] if t and isinstance(t[-1], list) else None
        if v is None or len(v) != 3 or not all(isinstance(x, str) and _is_number(x) for x in v):
            raise FoamDictError(f'keyword {key} in "{self.path()}": expected a vector, got {self._tokens(key)}')
        return [float(x) for x in v]


def _is_number(s: str) -> bool:
    try:
        float(s)
        return True
    except ValueError:
        return False


def _strip_comments(text: str) -> str:
    text = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)
    return re.sub(r"//[^\n]*", " ", text)


def parse(text: str, name: str = "") -> FoamDict:
    text = _strip_comments(text)
    for bad in ("#include", "#calc", "#codeStream", "#eval"):
        if bad in text:
            raise FoamDictError(f"{name}: {bad} directives are not supported")
    it = _Tokens(text)
    root = FoamDict(name)
    _parse_dict(it, root, top=True)
    root.pop("FoamFile", None)
    return root


class _Tokens:
    """token stream that remembers whether a token is glued to the previous one (no white space in between): OpenFOAM's
    keyType reads `div(phiJm,U)`, `interpolate(rho)`, `laplacian(taubyrhof,p)` as ONE keyword"""

    def __init__(self, text: str):
        self.m = list(_TOKEN.finditer(text))
        self.i = 0

    def __iter__(self):
        return self

    def __next__(self) -> str:
        if self.i >= len(self.m):
            raise StopIteration
        self.i += 1
        return self.m[self.i - 1].group(0)

    def glued_paren_follows(self) -> bool:
        return 0 < self.i < len(self.m) and self.m[self.i].group(0) == "(" and self.m[self.i].start() == self.m[self.i - 1].end()

    def take_balanced(self) -> str:
        """consume `( ... )` with nested parentheses and return its text without white space"""
        depth, out = 0, []
        for t in self:
            out.append(t)
            depth += (t == "(") - (t == ")")
            if depth == 0:
                return "".join(out)
        raise FoamDictError("unbalanced ( in a keyword")


def _parse_list(it: Iterator[str], close: str) -> List[Any]:
    out: List[Any] = ["["] if close == "]" else []
    for t in it:
        if t == close:
            return out
        if t == "(":
            out.append(_parse_list(it, ")"))
        elif t == "[":
            out.append(_parse_list(it, "]"))
        else:
            out.append(t)
    raise FoamDictError("unbalanced list")


def _expand(d: FoamDict, tok: str):
    """$name: the entry of this or an enclosing dictionary"""
    key, scope = tok[1:], d
    while scope is not None:
        v = scope._find(key)
        if v is not None:
            return v
        scope = scope.parent
    raise FoamDictError(f"macro {tok} is undefined")


def _parse_dict(it: Iterator[str], d: FoamDict, top: bool = False) -> None:
    for key in it:
        if key == "}":
            if top:
                raise FoamDictError("unbalanced }")
            return
        if key in ";":
            continue
        if isinstance(it, _Tokens) and not key.startswith('"') and it.glued_paren_follows():
            key += it.take_balanced()                                     # keyword with an argument list: div(phiJm,U)
        vals: List[Any] = []
        for t in it:
            if t == "{":
                if vals:
                    raise FoamDictError(f"entry {key}: unexpected {{ after values")
                sub = FoamDict(key, d)
                _parse_dict(it, sub)
                d[key] = sub
                break
            if t == ";":
                if len(vals) == 1 and isinstance(vals[0], FoamDict):      # `key $otherDict;`
                    d[key] = vals[0]
                else:
                    d[key] = vals
                break
            if t == "(":
                lst = _parse_list(it, ")")
                if vals and isinstance(vals[-1], str) and vals[-1].isdigit() and len(lst) == int(vals[-1]):
                    vals.pop()                                            # size prefix `3 ( a b c )`
                vals.append(lst)
            elif t == "[":
                vals.append(_parse_list(it, "]"))
            elif t.startswith("$"):
                v = _expand(d, t)
                if isinstance(v, FoamDict):
                    vals.append(v)
                else:
                    vals.extend(v)
            else:
                vals.append(t)
        else:
            raise FoamDictError(f"entry {key}: missing ; or }}")
    if not top:
        raise FoamDictError(f"dictionary {d.name}: missing }}")


def read(path: str) -> FoamDict:
    import os
    with open(path) as f:
        return parse(f.read(), os.path.basename(path))
