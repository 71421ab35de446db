"""basedet/structures/container.py:5-16 -- attribute dict whose indexing indexes every field."""


class Container(dict):
    def __init__(self, **kwargs):
        super().__init__(**kwargs)
        self.__dict__.update(kwargs)

    def __setattr__(self, key, value):
        super().__setattr__(key, value)
        dict.__setitem__(self, key, value)

    def __setitem__(self, key, value):
        dict.__setitem__(self, key, value)
        super().__setattr__(key, value)

    def __getitem__(self, idx):
        if isinstance(idx, str):
            return dict.__getitem__(self, idx)
        values = {}
        for k, v in vars(self).items():
            values[k] = v[idx]
        return Container(**values)

    def __str__(self):
        s = self.__class__.__name__ + "("
        s += "data=[{}])".format(", ".join((f"{k}: {v}" for k, v in vars(self).items())))
        return s
