"""BoundaryConditions / InternalConditions / Conditions with the reference's semantics
(/root/reference/src/conditions.jl), flattened to per-node and per-boundary-edge arrays instead
of Dicts."""
import numpy as np

Neumann, Dudt, Dirichlet, Constrained = "Neumann", "Dudt", "Dirichlet", "Constrained"
NODE_FREE, NODE_DIRICHLET, NODE_DUDT = 0, 1, 2
EDGE_NONE, EDGE_NEUMANN, EDGE_CONSTRAINED = 0, 1, 2


class BoundaryConditions:
    """conditions.jl:164-172,237-251: one function and one ConditionType per boundary section."""

    def __init__(self, mesh, functions, condition_types, parameters=None):
        if not isinstance(functions, (tuple, list)):
            functions = (functions,)
        if isinstance(condition_types, str):
            condition_types = (condition_types,)
        nsec = mesh.triangulation.num_sections
        if not (len(functions) == len(condition_types) == nsec):
            raise AssertionError("The number of boundary conditions must match the number of boundary sections (%d)." % nsec)
        for t in condition_types:
            if t not in (Neumann, Dudt, Dirichlet, Constrained):
                raise AssertionError("unknown condition type %r" % (t,))
        self.functions = tuple(functions)
        self.condition_types = tuple(condition_types)
        self.parameters = parameters  # kept for signature parity; registry specs carry their parameters

    def __repr__(self):  # Base.show, conditions.jl:173-180
        n = len(self.functions)
        if n > 1:
            return "BoundaryConditions with %d boundary conditions with types (%s)" % (n, ", ".join(self.condition_types))
        return "BoundaryConditions with %d boundary condition with type %s" % (n, self.condition_types[0])


class InternalConditions:
    """conditions.jl:223-230,253-269: dirichlet_nodes / dudt_nodes map node -> function index."""

    def __init__(self, functions=(), dirichlet_nodes=None, dudt_nodes=None, parameters=None):
        if not isinstance(functions, (tuple, list)):
            functions = (functions,)
        self.functions = tuple(functions)
        self.dirichlet_nodes = dict(dirichlet_nodes or {})
        self.dudt_nodes = dict(dudt_nodes or {})
        self.parameters = parameters

    def __repr__(self):  # Base.show, conditions.jl:231-235
        return "InternalConditions with %d Dirichlet nodes and %d Dudt nodes" % (len(self.dirichlet_nodes), len(self.dudt_nodes))


class Conditions:
    """conditions.jl:310-324 + merge_conditions! (:506-544).

    node_kind/node_fidx (N,) and edge_kind/edge_fidx (Eb,) follow the reference's rules:
    internal functions come first (fidx = section + nif), internal entries are overwritten by
    boundary entries, Dirichlet beats Dudt on the same node (source_contributions.jl:5-12)."""

    def __init__(self, mesh, bc: BoundaryConditions, ic: InternalConditions = None):
        ic = ic or InternalConditions()
        tri = mesh.triangulation
        N = tri.num_points
        nif = len(ic.functions)
        self.nif, self.condition_types, self.internal = nif, tuple(bc.condition_types), ic
        self.functions = tuple(ic.functions) + tuple(bc.functions)
        dir_f = np.full(N, -1, dtype=np.int32)
        dudt_f = np.full(N, -1, dtype=np.int32)
        for n, f in ic.dirichlet_nodes.items():
            dir_f[n] = f
        for n, f in ic.dudt_nodes.items():
            dudt_f[n] = f
        uv, sec = tri.boundary_edges()
        self.boundary_edges = uv
        self.edge_kind = np.zeros(len(uv), dtype=np.uint8)
        self.edge_fidx = np.zeros(len(uv), dtype=np.int32)
        for s, ctype in enumerate(bc.condition_types):
            sel = sec == s
            fidx = s + nif
            if ctype == Neumann:
                self.edge_kind[sel] = EDGE_NEUMANN
                self.edge_fidx[sel] = fidx
            elif ctype == Constrained:
                self.edge_kind[sel] = EDGE_CONSTRAINED
                self.edge_fidx[sel] = fidx
            elif ctype == Dirichlet:
                dir_f[uv[sel].ravel()] = fidx
            else:
                dudt_f[uv[sel].ravel()] = fidx
        self.dirichlet_fidx, self.dudt_fidx = dir_f, dudt_f
        self.node_kind = np.zeros(N, dtype=np.uint8)
        self.node_fidx = np.zeros(N, dtype=np.int32)
        is_dudt = dudt_f >= 0
        is_dir = dir_f >= 0
        self.node_kind[is_dudt] = NODE_DUDT
        self.node_fidx[is_dudt] = dudt_f[is_dudt]
        self.node_kind[is_dir] = NODE_DIRICHLET  # Dirichlet takes precedence
        self.node_fidx[is_dir] = dir_f[is_dir]

    def __repr__(self):  # Base.show, conditions.jl:325-335
        return "Conditions with\n   %d Neumann edges\n   %d Constrained edges\n   %d Dirichlet nodes\n   %d Dudt nodes" % (
            int((self.edge_kind == EDGE_NEUMANN).sum()), int((self.edge_kind == EDGE_CONSTRAINED).sum()),
            int((self.dirichlet_fidx >= 0).sum()), int((self.dudt_fidx >= 0).sum()))

    # the reference's predicates (conditions.jl:342-484)
    def is_dirichlet_node(self, i):
        return self.dirichlet_fidx[i] >= 0

    def is_dudt_node(self, i):
        return self.dudt_fidx[i] >= 0

    def has_condition(self, i):
        return self.node_kind[i] != NODE_FREE

    def has_dirichlet_nodes(self):
        return bool((self.dirichlet_fidx >= 0).any())

    def has_dudt_nodes(self):
        return bool((self.dudt_fidx >= 0).any())

    def has_constrained_edges(self):
        return bool((self.edge_kind == EDGE_CONSTRAINED).any())

    def has_neumann_edges(self):
        return bool((self.edge_kind == EDGE_NEUMANN).any())

    def get_dirichlet_nodes(self):
        idx = np.nonzero(self.dirichlet_fidx >= 0)[0]
        return dict(zip(idx.tolist(), self.dirichlet_fidx[idx].tolist()))

    def get_dudt_nodes(self):
        idx = np.nonzero(self.dudt_fidx >= 0)[0]
        return dict(zip(idx.tolist(), self.dudt_fidx[idx].tolist()))

    def get_neumann_edges(self):
        sel = self.edge_kind == EDGE_NEUMANN
        return {tuple(e): int(f) for e, f in zip(self.boundary_edges[sel].tolist(), self.edge_fidx[sel])}
