"""``StructureLibrary`` -- structures with their orientation lists, input of the OLD library api
(behavioural mirror of diffsims/libraries/structure_library.py:24-127)."""

__all__ = ["StructureLibrary"]


class StructureLibrary:
    """``struct_lib[identifier] = (structure, orientations)``; orientations are Euler triples (rzxz, degrees)."""

    def __init__(self, identifiers, structures, orientations):
        for what, seq in (("structures", structures), ("orientations", orientations)):
            if len(seq) != len(identifiers):
                raise ValueError(f"Number of identifiers ({len(identifiers)}) and {what} ({len(seq)}) "
                                 "must be the same.")
        self.identifiers, self.structures, self.orientations = identifiers, structures, orientations
        self.struct_lib = {name: (structure, rotations)
                           for name, structure, rotations in zip(identifiers, structures, orientations)}

    @classmethod
    def from_orientation_lists(cls, identifiers, structures, orientations):
        return cls(identifiers, structures, orientations)

    def get_library_size(self, to_print=False):
        """Total number of (structure, orientation) entries; optionally print the per-phase counts."""
        sizes = [len(rotations) for rotations in self.orientations]
        if to_print:
            for name, n in zip(self.identifiers, sizes):
                print(name, "has", n, "number of entries.")
            print("\nIn total:", sum(sizes), "number of entries")
        return sum(sizes)
