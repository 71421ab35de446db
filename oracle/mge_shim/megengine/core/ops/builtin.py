"""The handful of builtin ops that basedet/structures/op_patch.py builds its subgraphs from."""
import enum


class Elemwise:
    class Mode(enum.Enum):
        ADD = "add"
        SUB = "sub"
        MUL = "mul"
        TRUE_DIV = "true_div"
        MAX = "max"
        MIN = "min"
        POW = "pow"

    def __init__(self, mode):
        self.mode = mode


class AddAxis:
    def __init__(self, axis):
        self.axis = axis


class RemoveAxis:
    def __init__(self, axis):
        self.axis = axis


class Subtensor:
    def __init__(self, items):
        self.items = items


class Reduce:
    def __init__(self, mode="sum", axis=None):
        self.mode = mode
        self.axis = axis
