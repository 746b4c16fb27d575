"""flax.struct stand-in: frozen dataclasses with .replace (pytree registration is meaningless without tracing)."""
import dataclasses


def field(pytree_node=True, **kwargs):
    return dataclasses.field(metadata={"pytree_node": pytree_node}, **kwargs)


def dataclass(clz=None, **kw):
    if clz is None:
        return lambda c: dataclass(c, **kw)
    data_clz = dataclasses.dataclass(frozen=True)(clz)

    def replace(self, **updates):
        return dataclasses.replace(self, **updates)
    data_clz.replace = replace
    return data_clz
