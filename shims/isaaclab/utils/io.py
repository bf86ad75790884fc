import os
import pickle


def dump_yaml(filename, data, sort_keys=False):
    import yaml
    if not filename.endswith("yaml"):
        filename += ".yaml"
    os.makedirs(os.path.dirname(filename) or ".", exist_ok=True)
    if hasattr(data, "to_dict"):
        data = data.to_dict()
    with open(filename, "w") as f:
        yaml.dump(data, f, default_flow_style=False, sort_keys=sort_keys, Dumper=_Dumper())


def _Dumper():
    import yaml

    class D(yaml.SafeDumper):
        pass
    D.add_multi_representer(object, lambda dumper, obj: dumper.represent_str(str(obj)))
    return D


def dump_pickle(filename, data):
    if not filename.endswith("pkl"):
        filename += ".pkl"
    os.makedirs(os.path.dirname(filename) or ".", exist_ok=True)
    with open(filename, "wb") as f:
        pickle.dump(data, f)


def load_pickle(filename):
    with open(filename, "rb") as f:
        return pickle.load(f)
