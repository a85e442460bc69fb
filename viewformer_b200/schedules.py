"""Scalar schedules written as strings in the reference's configs (viewformer/utils/schedules.py:72-247): the value of
``localization_weight`` (models/config.py) as a function of the training step (migt.py:268, 446).

Forms understood: a number ("1", "0.5"), ``linear(a,b[,n])``, ``cosine(a,b[,n])`` and ``warmup(<schedule>,steps)``; ``n`` (the horizon)
may be left out and filled in later with ``with_total_steps`` (migt.py:268 does that with ``config.total_steps``).
  linear:  a + min(t / n, 1) (b - a)                                   (schedules.py:174-175)
  cosine:  b + (a - b) / 2 * (cos(pi * min(1, t / n)) + 1)             (schedules.py:199-201)
  warmup:  min(t, w) / w * inner(max(t - w, 0))                        (schedules.py:223-226)
Everything is evaluated on the host in Python floats: one number per optimisation step.

Note on the reference: its generic string parser (`default_from_str`, schedules.py:6-21) builds the positional-argument type list from
``{POSITIONAL_ONLY or POSITIONAL_OR_KEYWORD}`` (= ``{POSITIONAL_ONLY}``), so ``linear(...)`` / ``cosine(...)`` strings trip its own
assertion; only numbers and ``warmup(<number>,steps)`` parse there.  The forms above follow what the strings say.
"""
import math


class Schedule:
    def __call__(self, t):
        raise NotImplementedError

    def with_total_steps(self, n):
        return self

    def is_zero(self):
        return False

    @staticmethod
    def from_str(text):
        return parse(text)


class Constant(Schedule):
    def __init__(self, value):
        self.value = float(value)

    def __call__(self, t):
        return self.value

    def is_zero(self):
        return self.value == 0.0

    def __str__(self):
        return repr(self.value)


class _Ramp(Schedule):
    name = ""

    def __init__(self, initial_value, final_value, num_total_steps=None):
        self.a, self.b = float(initial_value), float(final_value)
        self.n = None if num_total_steps is None else int(num_total_steps)

    def with_total_steps(self, n):
        return self if self.n is not None else type(self)(self.a, self.b, n)

    def is_zero(self):
        return self.a == 0.0 and self.b == 0.0

    def _frac(self, t):
        if self.n is None:
            raise ValueError(f"{self}: the schedule has no horizon; call with_total_steps(config.total_steps) first")
        return min(float(t) / self.n, 1.0)

    def __str__(self):
        return f"{self.name}({self.a},{self.b},{self.n})"


class Linear(_Ramp):
    name = "linear"

    def __call__(self, t):
        return self.a + self._frac(t) * (self.b - self.a)


class Cosine(_Ramp):
    name = "cosine"

    def __call__(self, t):
        return self.b + (self.a - self.b) * 0.5 * (math.cos(self._frac(t) * math.pi) + 1.0)


class Warmup(Schedule):
    def __init__(self, inner, warmup_steps):
        self.inner, self.w = inner, int(warmup_steps)

    def with_total_steps(self, n):
        return Warmup(self.inner.with_total_steps(n), self.w)

    def is_zero(self):
        return self.inner.is_zero()

    def __call__(self, t):
        t = float(t)
        return (min(t, self.w) / self.w) * self.inner(max(t - self.w, 0.0))

    def __str__(self):
        return f"warmup({self.inner},{self.w})"


def parse(text):
    """String (or number, or Schedule) -> Schedule."""
    if isinstance(text, Schedule):
        return text
    if isinstance(text, (int, float)):
        return Constant(text)
    s = str(text).strip()
    try:
        return Constant(float(s))
    except ValueError:
        pass
    if s.startswith("warmup(") and s.endswith(")") and "," in s:
        body = s[len("warmup("):-1]
        cut = body.rindex(",")
        return Warmup(parse(body[:cut]), int(body[cut + 1:].strip()))
    for cls in (Linear, Cosine):
        head = cls.name + "("
        if s.startswith(head) and s.endswith(")"):
            args = [a.strip() for a in s[len(head):-1].split(",") if a.strip()]
            kw = dict(a.split("=", 1) for a in args if "=" in a)
            pos = [a for a in args if "=" not in a]
            names = ["initial_value", "final_value", "num_total_steps"]
            vals = dict(zip(names, pos))
            vals.update({k.strip(): v.strip() for k, v in kw.items()})
            if "initial_value" not in vals or "final_value" not in vals or set(vals) - set(names):
                raise ValueError(f"cannot parse schedule '{text}'")
            n = vals.get("num_total_steps")
            return cls(float(vals["initial_value"]), float(vals["final_value"]), None if n in (None, "None") else int(float(n)))
    raise ValueError(f"cannot parse schedule '{text}'")
