"""Result records of the bulk computations, with the attribute names of the reference's
`coffe.representation` (python/coffe/representation.py:47-330): read-only properties, `to_dict()` for
`pandas.DataFrame([_.to_dict() for _ in result])`."""


class Representation:
    _fields = ()

    def __init__(self, **kw):
        missing = [f for f in self._fields if f not in kw]
        extra = [k for k in kw if k not in self._fields]
        if missing or extra:
            raise TypeError(f"{type(self).__name__} takes exactly the keyword arguments {self._fields}")
        self._values = {f: kw[f] for f in self._fields}

    def __getattr__(self, name):
        values = self.__dict__.get("_values", {})
        if name in values:
            return values[name]
        raise AttributeError(name)

    def __setattr__(self, name, value):
        if name in self._fields:
            raise AttributeError(f"{name} is read-only")
        super().__setattr__(name, value)

    def to_dict(self):
        return dict(self._values)

    def __repr__(self):
        return f"<{type(self).__name__}({self.to_dict()})>"

    def __str__(self):
        return f"{type(self).__name__}({self.to_dict()})"

    def __eq__(self, other):
        return type(other) is type(self) and other._values == self._values


class Corrfunc(Representation):
    _fields = ("r", "mu", "z", "value")


class Multipoles(Representation):
    _fields = ("l", "r", "z", "value")


class AverageMultipoles(Representation):
    _fields = ("l", "r", "z_min", "z_max", "value")


class Covariance(Representation):
    _fields = ("r1", "r2", "l1", "l2", "z", "value")


class AverageCovariance(Representation):
    _fields = ("r1", "r2", "l1", "l2", "z_min", "z_max", "value")
