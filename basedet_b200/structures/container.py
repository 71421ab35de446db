"""basedet/structures/container.py:5-16 -- attribute dict whose indexing indexes every field."""


class Container(dict):
    """Fields are reachable as attributes and as keys; ``c[idx]`` with a non-string index slices every field."""

    def __init__(self, **fields):
        super().__init__()
        for name, value in fields.items():
            self[name] = value

    def __setattr__(self, name, value):
        self[name] = value

    def __setitem__(self, name, value):
        dict.__setitem__(self, name, value)
        object.__setattr__(self, name, value)

    def __getitem__(self, idx):
        if isinstance(idx, str):
            return dict.__getitem__(self, idx)
        return type(self)(**{name: value[idx] for name, value in self.items()})

    def __str__(self):
        body = ", ".join("{}: {}".format(name, value) for name, value in self.items())
        return "{}(data=[{}])".format(type(self).__name__, body)
