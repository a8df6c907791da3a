"""Path helper used by every ``write_to_file`` / ``from_file`` (same rule as /root/reference/amof/files/path.py:7-21)."""
import pathlib


def append_suffix(path, suffix):
    """Return ``path`` as a pathlib.Path whose last suffix is ``suffix`` (a leading '.' is added when missing);
    the suffix is appended only when it is not already the last one."""
    if suffix and not suffix.startswith('.'):
        suffix = '.' + suffix
    path = pathlib.Path(path)
    if path.suffix == suffix:
        return path
    return path.with_name(path.name + suffix)
