"""``StructureLibrary`` (diffsims/libraries/structure_library.py:24-127): structures + orientation lists."""

__all__ = ["StructureLibrary"]


class StructureLibrary:
    """Identifiers, structures and per-structure lists of Euler angles (rzxz, degrees)."""

    def __init__(self, identifiers, structures, orientations):
        if len(identifiers) != len(structures):
            raise ValueError("Number of identifiers ({}) and structures ({}) must be the same.".format(
                len(identifiers), len(structures)))
        if len(identifiers) != len(orientations):
            raise ValueError("Number of identifiers ({}) and orientations ({}) must be the same.".format(
                len(identifiers), len(orientations)))
        self.identifiers = identifiers
        self.structures = structures
        self.orientations = orientations
        self.struct_lib = dict()
        for ident, struct, ori in zip(identifiers, structures, orientations):
            self.struct_lib[ident] = (struct, ori)

    @classmethod
    def from_orientation_lists(cls, identifiers, structures, orientations):
        return cls(identifiers, structures, orientations)

    def get_library_size(self, to_print=False):
        size_library = 0
        for ident, ori in zip(self.identifiers, self.orientations):
            size_library += 1 if len(ori) == 1 else len(ori)
            if to_print:
                print(ident, "has", len(ori), "number of entries.")
        if to_print:
            print("\nIn total:", size_library, "number of entries")
        return size_library
