# FVMCuda.jl -- the reference-side binding a FiniteVolumeMethod.jl maintainer would add to make
# libfvmcuda.so a drop-in for `fvm_eqs!` and the template operators.
#
# UNTESTED IN THIS REPOSITORY: Julia is not installed in the authoring image.  The file documents the
# exact ccall signatures of include/fvmcuda.h and where they hook into the reference
# (src/solve.jl:1-42, src/equations/main_equations.jl:28-35, src/specific_problems/abstract_templates.jl:58-60).
module FVMCuda

using FiniteVolumeMethod
using DelaunayTriangulation
const FVM = FiniteVolumeMethod
const LIB = get(ENV, "FVMCUDA_LIB", "libfvmcuda.so")

struct FVMCudaError <: Exception
    code::Int32
    msg::String
end
check(h, rc) = rc == 0 || throw(FVMCudaError(rc, unsafe_string(ccall((:fvm_last_error, LIB), Cstring, (Ptr{Cvoid},), h))))

# registry specs (the device cannot run Julia closures; anything else must raise before any launch)
struct ConstantDiffusion; D::Float64; end
struct PowerDiffusion; D0::Float64; m::Float64; use_abs::Bool; end
struct Const; c::Float64; end
struct AffineU; c0::Float64; c1::Float64; end

mutable struct Handle
    ptr::Ptr{Cvoid}
    function Handle(tri::Triangulation, neq::Integer; device = 0)
        pts = collect(Float64, Iterators.flatten(DelaunayTriangulation.each_point(tri)))   # interleaved x,y
        T = collect(Int32, Iterators.flatten(triangle_vertices(t) for t in each_solid_triangle(tri)))
        out = Ref{Ptr{Cvoid}}(C_NULL)
        rc = ccall((:fvm_create, LIB), Int32,
            (Ptr{Float64}, Int64, Ptr{Int32}, Int64, Int32, Int32, Int32, Ptr{Ptr{Cvoid}}),
            pts, length(pts) ÷ 2, T, length(T) ÷ 3, 1 #= index_base: Julia is 1-based =#, neq, device, out)
        rc == 0 || throw(FVMCudaError(rc, unsafe_string(ccall((:fvm_last_error, LIB), Cstring, (Ptr{Cvoid},), C_NULL))))
        h = new(out[])
        finalizer(x -> ccall((:fvm_destroy, LIB), Int32, (Ptr{Cvoid},), x.ptr), h)
        return h
    end
end

"Sibling of `get_multithreading_parameters` (src/solve.jl:1-27): flattens `prob` into the handle once."
function get_cuda_parameters(prob::FVMProblem; tile_triangles = 0, geometry_mode = 0)
    tri = prob.mesh.triangulation
    h = Handle(tri, 1)
    edges = collect(keys(get_boundary_edge_map(tri)))
    uv = collect(Int32, Iterators.flatten(edges))
    check(h.ptr, ccall((:fvm_set_boundary_edges, LIB), Int32, (Ptr{Cvoid}, Ptr{Int32}, Int64), h.ptr, uv, length(edges)))
    conds = prob.conditions
    N = DelaunayTriangulation.num_points(tri)
    nkind = zeros(UInt8, N); nfidx = zeros(Int32, N)
    for (i, f) in conds.dudt_nodes;      nkind[i] = 2; nfidx[i] = f - 1; end
    for (i, f) in conds.dirichlet_nodes; nkind[i] = 1; nfidx[i] = f - 1; end   # Dirichlet beats Dudt
    ekind = [haskey(conds.neumann_edges, e) ? UInt8(1) : haskey(conds.constrained_edges, e) ? UInt8(2) : UInt8(0) for e in edges]
    efidx = Int32[get(conds.neumann_edges, e, get(conds.constrained_edges, e, 1)) - 1 for e in edges]
    check(h.ptr, ccall((:fvm_set_edge_conditions, LIB), Int32, (Ptr{Cvoid}, Int32, Ptr{UInt8}, Ptr{Int32}), h.ptr, 0, ekind, efidx))
    check(h.ptr, ccall((:fvm_set_node_conditions, LIB), Int32, (Ptr{Cvoid}, Int32, Ptr{UInt8}, Ptr{Int32}), h.ptr, 0, nkind, nfidx))
    for (fidx, f) in enumerate(conds.functions)     # f.fnc must be a registry spec, e.g. Const(0.0)
        id, p = f.fnc isa Const ? (0, [f.fnc.c]) : f.fnc isa AffineU ? (1, [f.fnc.c0, f.fnc.c1]) :
            throw(ArgumentError("condition function $(f.fnc) is not in the compiled device registry"))
        check(h.ptr, ccall((:fvm_set_condition_fn, LIB), Int32, (Ptr{Cvoid}, Int32, Int32, Int32, Ptr{Float64}, Int32),
            h.ptr, 0, fidx - 1, id, p, length(p)))
    end
    D = prob.flux_function      # a registry spec stored instead of the closure of construct_flux_function
    model, p = D isa ConstantDiffusion ? (0, [D.D]) : D isa PowerDiffusion ? (2, [D.D0, D.m, Float64(D.use_abs)]) :
        throw(ArgumentError("flux function $(D) is not in the compiled device registry"))
    check(h.ptr, ccall((:fvm_set_flux, LIB), Int32, (Ptr{Cvoid}, Int32, Ptr{Float64}, Int32), h.ptr, model, p, length(p)))
    check(h.ptr, ccall((:fvm_finalize, LIB), Int32, (Ptr{Cvoid}, Int32, Int32), h.ptr, tile_triangles, geometry_mode))
    return (prob = prob, parallel = Val(:cuda), handle = h)
end

"`fvm_eqs!` method for `p.parallel == Val(:cuda)` (src/equations/main_equations.jl:28-35)."
function fvm_eqs!(du::Array{Float64}, u::Array{Float64}, p::NamedTuple{(:prob, :parallel, :handle)}, t)
    check(p.handle.ptr, ccall((:fvm_rhs, LIB), Int32, (Ptr{Cvoid}, Float64, Ptr{Float64}, Ptr{Float64}, Int32),
        p.handle.ptr, Float64(t), u, du, 0))
    return du
end

"Dirichlet callback body (src/equations/dirichlet.jl:78-86)."
function update_dirichlet_nodes!(integrator)
    p = integrator.p
    check(p.handle.ptr, ccall((:fvm_apply_dirichlet, LIB), Int32, (Ptr{Cvoid}, Float64, Ptr{Float64}, Int32),
        p.handle.ptr, Float64(integrator.t), integrator.u, 0))
    return nothing
end

"Page-locks the arrays an integrator passes to `fvm_eqs!` / `mul!` (full PCIe rate, pipelined copies); undo with `unpin!`."
pin!(a::Array{Float64}) = (ccall((:fvm_host_register, LIB), Int32, (Ptr{Cvoid}, Int64), a, sizeof(a)) == 0 || error("fvm_host_register failed"); a)
unpin!(a::Array{Float64}) = (ccall((:fvm_host_unregister, LIB), Int32, (Ptr{Cvoid},), a); a)

"Template operator: `mul!(du, A, u)` of the MatrixOperator (diffusion_equation.jl:93-94)."
struct FVMCudaOperator; handle::Handle; n::Int; end
function LinearAlgebra_mul!(du::Vector{Float64}, A::FVMCudaOperator, u::Vector{Float64})
    check(A.handle.ptr, ccall((:fvm_spmv, LIB), Int32, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Int32, Int32),
        A.handle.ptr, u, du, 1, 0))
    return du
end

"Device-resident `solve(prob, Tsit5(); adaptive = false, dt)`."
function tsit5!(u::Array{Float64}, h::Handle, t0, t1, dt; use_operator = false)
    check(h.ptr, ccall((:fvm_tsit5, LIB), Int32,
        (Ptr{Cvoid}, Int32, Ptr{Float64}, Float64, Float64, Float64, Int64, Ptr{Float64}, Ptr{Float64}, Int32),
        h.ptr, use_operator, u, t0, t1, dt, 0, C_NULL, C_NULL, 0))
    return u
end

# ---- FVMWIRE containers: the importer/exporter for DelaunayTriangulation objects and solutions -------
wire_check(w, rc) = rc == 0 || throw(FVMCudaError(rc, unsafe_string(ccall((:fvm_wire_last_error, LIB), Cstring, (Ptr{Cvoid},), w))))
const WIRE_DTYPE = Dict(Float64 => Int32(1), Int32 => Int32(2), UInt8 => Int32(3), Int64 => Int32(4))

function wire_put(w::Ptr{Cvoid}, name::String, A::Array{T}) where {T}
    dims = collect(Int64, size(A))          # Julia's size(A) is already fastest-first
    wire_check(w, ccall((:fvm_wire_put, LIB), Int32, (Ptr{Cvoid}, Cstring, Int32, Int32, Ptr{Int64}, Ptr{Cvoid}),
        w, name, WIRE_DTYPE[T], length(dims), dims, A))
end

"Writes the arrays `FVMGeometry(tri)` reads (src/geometry.jl:99-106) plus the boundary description."
function write_mesh(path::String, tri::Triangulation)
    out = Ref{Ptr{Cvoid}}(C_NULL)
    wire_check(C_NULL, ccall((:fvm_wire_create, LIB), Int32, (Cstring, Ptr{Ptr{Cvoid}}), path, out))
    w = out[]
    pts = reshape(collect(Float64, Iterators.flatten(DelaunayTriangulation.each_point(tri))), 2, :)
    T = reshape(collect(Int32, Iterators.flatten(triangle_vertices(t) for t in each_solid_triangle(tri))), 3, :)
    edges = collect(keys(get_boundary_edge_map(tri)))
    wire_put(w, "points", pts); wire_put(w, "triangles", T); wire_put(w, "index_base", Int32[1])
    wire_put(w, "boundary_edges", reshape(collect(Int32, Iterators.flatten(edges)), 2, :))
    # section of an edge (u, v): the ghost vertex on its other side is -section (src/conditions.jl:507-515)
    wire_put(w, "boundary_edge_section", Int32[-get_adjacent(tri, v, u) - 1 for (u, v) in edges])
    wire_put(w, "num_sections", Int32[length(DelaunayTriangulation.get_ghost_vertex_map(tri))])
    wire_check(C_NULL, ccall((:fvm_wire_close, LIB), Int32, (Ptr{Cvoid},), w))
    return path
end

"`sol.u` / `sol.t` of `solve(prob, alg; saveat)` (src/solve.jl:197-208) as one (N, nsave) or (neq, N, nsave) array."
function write_solution(path::String, sol)
    out = Ref{Ptr{Cvoid}}(C_NULL)
    wire_check(C_NULL, ccall((:fvm_wire_create, LIB), Int32, (Cstring, Ptr{Ptr{Cvoid}}), path, out))
    w = out[]
    wire_put(w, "u", cat(sol.u...; dims = ndims(sol.u[1]) + 1)); wire_put(w, "t", collect(Float64, sol.t))
    wire_check(C_NULL, ccall((:fvm_wire_close, LIB), Int32, (Ptr{Cvoid},), w))
    return path
end

"`Handle` straight from a mesh container (no Triangulation object on the Julia side)."
function handle_from_wire(path::String, neq::Integer; device = 0)
    out = Ref{Ptr{Cvoid}}(C_NULL)
    rc = ccall((:fvm_create_from_wire, LIB), Int32, (Cstring, Int32, Int32, Ptr{Ptr{Cvoid}}), path, neq, device, out)
    rc == 0 || throw(FVMCudaError(rc, unsafe_string(ccall((:fvm_wire_last_error, LIB), Cstring, (Ptr{Cvoid},), C_NULL))))
    return out[]
end

end # module
