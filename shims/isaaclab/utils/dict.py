def print_dict(val, nesting: int = -4, start: bool = True):
    if isinstance(val, dict):
        if not start:
            print("")
        nesting += 4
        for k in val:
            print(nesting * " ", end="")
            print(k, end=": ")
            print_dict(val[k], nesting, start=False)
    else:
        print(val)


def update_class_from_dict(obj, data, _ns: str = ""):
    for key, value in dict(data).items():
        if not hasattr(obj, key):
            raise KeyError(f"[Config]: Key not found under namespace: {_ns}/{key}")
        cur = getattr(obj, key)
        if hasattr(value, "items") and not isinstance(cur, dict) and hasattr(cur, "__dict__"):
            update_class_from_dict(cur, value, _ns + "/" + key)
        else:
            if isinstance(cur, tuple) and isinstance(value, list):
                value = tuple(value)
            setattr(obj, key, value)
