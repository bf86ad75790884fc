from .configclass import configclass  # noqa: F401


def replace_slices_with_strings(data):
    if isinstance(data, dict):
        return {k: replace_slices_with_strings(v) for k, v in data.items()}
    if isinstance(data, slice):
        return f"slice({data.start},{data.stop},{data.step})"
    return data


def replace_strings_with_slices(data):
    if isinstance(data, dict):
        return {k: replace_strings_with_slices(v) for k, v in data.items()}
    if isinstance(data, str) and data.startswith("slice("):
        a = [None if x == "None" else int(x) for x in data[6:-1].split(",")]
        return slice(*a)
    return data
