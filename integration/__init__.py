"""Reference-side adapters: libsbx behind the reference's own seams B3 and B2."""
