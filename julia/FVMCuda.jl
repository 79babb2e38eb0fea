# FVMCuda.jl -- the reference-side binding a FiniteVolumeMethod.jl maintainer adds to make libfvmcuda.so a
# drop-in for `fvm_eqs!`, the Dirichlet callback, the template operators and their solves.
#
# Julia is not installed in the authoring image, so this file cannot be executed here.  What CAN be checked is
# checked: tests/test_julia_binding_cpu.py parses every `ccall` below and verifies its name, arity, argument
# types and return type against include/fvmcuda.h, and that every setter of the ABI is reached.
#
# Hooks into the reference (file:line relative to FiniteVolumeMethod.jl v1.2.3):
#   get_cuda_parameters(prob)           sibling of get_multithreading_parameters      src/solve.jl:1-27
#   fvm_eqs!(du, u, p, t)               method for p.parallel == Val(:cuda)           src/equations/main_equations.jl:28-35
#   update_dirichlet_nodes!(integrator)                                              src/equations/dirichlet.jl:78-86
#   cuda_jac_prototype / fvm_jacobian!  jac_prototype of the ODEFunction              src/solve.jl:50-131,167-183
#   FVMCudaOperator + mul!              MatrixOperator(sparse(Afull))                 src/specific_problems/diffusion_equation.jl:82-94
#   solve(prob, FVMCudaKrylov())        solve(prob::AbstractFVMTemplate, alg)         src/specific_problems/abstract_templates.jl:58-60
module FVMCuda

using FiniteVolumeMethod
using DelaunayTriangulation
using LinearAlgebra
using SparseArrays
import CommonSolve
const FVM = FiniteVolumeMethod
const DT = DelaunayTriangulation
const LIB = get(ENV, "FVMCUDA_LIB", "libfvmcuda.so")

struct FVMCudaError <: Exception
    code::Int32
    msg::String
end
last_error(h) = unsafe_string(ccall((:fvm_last_error, LIB), Cstring, (Ptr{Cvoid},), h))
check(h, rc) = rc == 0 || throw(rc == 3 ? ArgumentError(last_error(h)) : FVMCudaError(rc, last_error(h)))

# ---------------------------------------------------------------------------------------------------------
# The device registry.  Julia closures cannot run on the GPU: a flux / diffusion / source / condition function
# must be one of these callable structs (they are ordinary functions for the reference's CPU path too, so one
# problem definition runs on both).  Anything else raises ArgumentError before any kernel launch.
# ---------------------------------------------------------------------------------------------------------
abstract type DeviceFunctor <: Function end

# diffusion functions D(x, y, t, u, p) (src/problem.jl:425-440) and flux functions q(x, y, t, α, β, γ, p)
struct ConstantDiffusion <: DeviceFunctor; D::Float64; end
(f::ConstantDiffusion)(x, y, t, u, p) = f.D
struct TabulatedDiffusion{F} <: DeviceFunctor; fn::F; end            # any D(x, y): tabulated on the host at setup
(f::TabulatedDiffusion)(x, y, t, u, p) = f.fn(x, y)
struct PowerDiffusion <: DeviceFunctor; D0::Float64; m::Float64; use_abs::Bool; end
PowerDiffusion(D0, m) = PowerDiffusion(D0, m, false)
(f::PowerDiffusion)(x, y, t, u, p) = f.D0 * (f.use_abs ? abs(u) : u)^(f.m - 1)
struct AdvectionDiffusionFlux <: DeviceFunctor; D::Float64; nu_x::Float64; nu_y::Float64; end
function (f::AdvectionDiffusionFlux)(x, y, t, α, β, γ, p)
    u = α * x + β * y + γ
    return (f.nu_x * u - f.D * α, f.nu_y * u - f.D * β)
end
struct KellerSegelFlux <: DeviceFunctor; c::Float64; D::Float64; var::Int; end   # src/FiniteVolumeMethod.jl:98-110
function (f::KellerSegelFlux)(x, y, t, α, β, γ, p)
    f.var == 2 && return (-f.D * α[2], -f.D * β[2])
    u = α[1] * x + β[1] * y + γ[1]
    χ = f.c * u / (1 + u^2)
    return (χ * α[2] - α[1], χ * β[2] - β[1])
end

# sources S(x, y, t, u, p) (src/problem.jl:10-13, 342-345); systems see the tuple of all species
struct ZeroSource <: DeviceFunctor end
(f::ZeroSource)(x, y, t, u, p) = zero(eltype(u))
struct LinearSource <: DeviceFunctor; lam::Float64; mu::Float64; var::Int; end
LinearSource(lam, mu = 0.0) = LinearSource(lam, mu, 0)
(f::LinearSource)(x, y, t, u, p) = f.lam * (f.var == 0 ? u : u[f.var]) + f.mu
struct LogisticSource <: DeviceFunctor; lam::Float64; var::Int; end
LogisticSource(lam) = LogisticSource(lam, 0)
(f::LogisticSource)(x, y, t, u, p) = (w = f.var == 0 ? u : u[f.var]; f.lam * w * (1 - w))
struct TabulatedSource{F} <: DeviceFunctor; fn::F; end               # any S(x, y)
(f::TabulatedSource)(x, y, t, u, p) = f.fn(x, y)
struct GrayScottSource <: DeviceFunctor; b::Float64; d::Float64; var::Int; end
(f::GrayScottSource)(x, y, t, u, p) = f.var == 1 ? f.b * (1 - u[1]) - u[1] * u[2]^2 : -f.d * u[2] + u[1] * u[2]^2
struct BrusselatorSource <: DeviceFunctor; var::Int; end
(f::BrusselatorSource)(x, y, t, u, p) = f.var == 1 ? u[1]^2 * u[2] - 2u[1] : -u[1]^2 * u[2] + u[1]
struct KellerSegelSource <: DeviceFunctor; a::Float64; var::Int; end
(f::KellerSegelSource)(x, y, t, u, p) = f.var == 1 ? u[1] * (1 - u[1]) : u[1] - f.a * u[2]

# boundary / internal condition functions a(x, y, t, u, p) (src/conditions.jl:16-20)
struct Const <: DeviceFunctor; c::Float64; end
(f::Const)(x, y, t, u, p) = f.c
struct AffineU <: DeviceFunctor; c0::Float64; c1::Float64; var::Int; end
AffineU(c0, c1) = AffineU(c0, c1, 0)
(f::AffineU)(x, y, t, u, p) = f.c0 + f.c1 * (f.var == 0 ? u : u[f.var])
struct ExpSaturation <: DeviceFunctor; c0::Float64; tau::Float64; end
(f::ExpSaturation)(x, y, t, u, p) = f.c0 * (1 - exp(-t / f.tau))
struct LinearXY <: DeviceFunctor; c0::Float64; cx::Float64; cy::Float64; end
(f::LinearXY)(x, y, t, u, p) = f.c0 + f.cx * x + f.cy * y
struct ExpXYT <: DeviceFunctor; c0::Float64; cx::Float64; cy::Float64; ct::Float64; end
(f::ExpXYT)(x, y, t, u, p) = f.c0 * exp(f.cx * x + f.cy * y + f.ct * t)

unsupported(kind, f) = throw(ArgumentError("$kind function $(f) is not in the compiled device registry " *
    "(arbitrary Julia closures cannot run on the GPU); use one of the FVMCuda functor structs"))

cond_spec(f::Const) = (0, [f.c])
cond_spec(f::AffineU) = (1, [f.c0, f.c1])
cond_spec(f::ExpSaturation) = (2, [f.c0, f.tau])
cond_spec(f::LinearXY) = (3, [f.c0, f.cx, f.cy])
cond_spec(f::ExpXYT) = (4, [f.c0, f.cx, f.cy, f.ct])
cond_spec(f) = unsupported("condition", f)

"The registry spec behind `prob.flux_function`: the spec itself, or the `D` captured by construct_flux_function (src/problem.jl:425-440)."
function flux_spec(f)
    f isa DeviceFunctor && return f
    if hasfield(typeof(f), :D) && getfield(f, :D) isa DeviceFunctor
        return getfield(f, :D)
    end
    unsupported("flux / diffusion", f)
end

# ---------------------------------------------------------------------------------------------------------
# handle + mesh flattening (FVMGeometry(tri), src/geometry.jl:99-169)
# ---------------------------------------------------------------------------------------------------------
mutable struct Handle
    ptr::Ptr{Cvoid}
    triangles::Vector{NTuple{3, Int}}      # solid triangles in the order handed to the library
    edges::Vector{NTuple{2, Int}}          # keys(get_boundary_edge_map(tri)) in the order handed to the library
    n::Int                                 # number of points
    neq::Int
end

function Handle(tri::Triangulation, neq::Integer; device = 0)
    pts = collect(Float64, Iterators.flatten(DT.each_point(tri)))                       # interleaved x, y
    tris = [Tuple(Int.(triangle_vertices(T))) for T in each_solid_triangle(tri)]
    T = collect(Int32, Iterators.flatten(tris))
    out = Ref{Ptr{Cvoid}}(C_NULL)
    rc = ccall((:fvm_create, LIB), Int32,
        (Ptr{Float64}, Int64, Ptr{Int32}, Int64, Int32, Int32, Int32, Ptr{Ptr{Cvoid}}),
        pts, length(pts) ÷ 2, T, length(tris), 1 #= index_base: Julia is 1-based =#, neq, device, out)
    rc == 0 || throw(FVMCudaError(rc, last_error(C_NULL)))
    edges = [Tuple(Int.(e)) for e in keys(get_boundary_edge_map(tri))]
    h = Handle(out[], tris, edges, length(pts) ÷ 2, neq)
    finalizer(x -> ccall((:fvm_destroy, LIB), Int32, (Ptr{Cvoid},), x.ptr), h)
    uv = collect(Int32, Iterators.flatten(edges))
    check(h.ptr, ccall((:fvm_set_boundary_edges, LIB), Int32, (Ptr{Cvoid}, Ptr{Int32}, Int64), h.ptr, uv, length(edges)))
    return h
end

"Flattens one species' `Conditions` (src/conditions.jl:310-324, 506-544) into the per-node / per-edge arrays of the ABI."
function set_conditions!(h::Handle, var::Integer, conds; register_functions = true)
    nkind = zeros(UInt8, h.n); nfidx = zeros(Int32, h.n)
    for (i, f) in conds.dudt_nodes;      nkind[i] = 2; nfidx[i] = f - 1; end
    for (i, f) in conds.dirichlet_nodes; nkind[i] = 1; nfidx[i] = f - 1; end           # Dirichlet beats Dudt (source_contributions.jl:5-12)
    ekind = UInt8[haskey(conds.neumann_edges, e) ? 1 : haskey(conds.constrained_edges, e) ? 2 : 0 for e in h.edges]
    efidx = Int32[get(conds.neumann_edges, e, get(conds.constrained_edges, e, 1)) - 1 for e in h.edges]
    check(h.ptr, ccall((:fvm_set_edge_conditions, LIB), Int32, (Ptr{Cvoid}, Int32, Ptr{UInt8}, Ptr{Int32}), h.ptr, var, ekind, efidx))
    check(h.ptr, ccall((:fvm_set_node_conditions, LIB), Int32, (Ptr{Cvoid}, Int32, Ptr{UInt8}, Ptr{Int32}), h.ptr, var, nkind, nfidx))
    register_functions || return nothing
    used = Set{Int}(vcat(collect(values(conds.dudt_nodes)), collect(values(conds.dirichlet_nodes)), collect(values(conds.neumann_edges))))
    for fidx in sort!(collect(used))                                                    # only functions the RHS can call
        id, p = cond_spec(conds.functions[fidx].fnc)
        check(h.ptr, ccall((:fvm_set_condition_fn, LIB), Int32, (Ptr{Cvoid}, Int32, Int32, Int32, Ptr{Float64}, Int32),
            h.ptr, var, fidx - 1, id, p, length(p)))
    end
    return nothing
end

"cv-edge midpoints of every solid triangle, [T][3], in the handle's triangle order (get_cv_components, src/geometry/control_volumes.jl:16-21)."
function cv_edge_midpoints(mesh::FVMGeometry, h::Handle)
    xs = Matrix{Float64}(undef, 3, length(h.triangles)); ys = similar(xs)
    for (k, T) in enumerate(h.triangles)
        props = mesh.triangle_props[T]
        for e in 1:3
            x, y, _, _, _ = FVM.get_cv_components(props, e)
            xs[e, k] = x; ys[e, k] = y
        end
    end
    return xs, ys
end
"the two quarter points of every boundary edge, [Eb][2] (get_boundary_cv_components, control_volumes.jl:41-56)."
function boundary_quarter_points(tri::Triangulation, h::Handle)
    xs = Matrix{Float64}(undef, 2, length(h.edges)); ys = similar(xs)
    for (k, (i, j)) in enumerate(h.edges)
        p, q = get_point(tri, i, j)
        px, py = getxy(p); qx, qy = getxy(q)
        mx, my = (px + qx) / 2, (py + qy) / 2
        xs[1, k] = (px + mx) / 2; ys[1, k] = (py + my) / 2
        xs[2, k] = (qx + mx) / 2; ys[2, k] = (qy + my) / 2
    end
    return xs, ys
end

function set_flux!(h::Handle, mesh::FVMGeometry, specs)
    s1 = specs[1]
    all(s -> typeof(s) === typeof(s1), specs) || throw(ArgumentError("all species of an FVMSystem must use the same flux model"))
    if s1 isa TabulatedDiffusion
        xs, ys = cv_edge_midpoints(mesh, h); bx, by = boundary_quarter_points(mesh.triangulation, h)
        dcv = Float64[s1.fn(xs[e, k], ys[e, k]) for e in 1:3, k in axes(xs, 2)]            # column-major (3, T) == C [T][3]
        dbn = Float64[s1.fn(bx[e, k], by[e, k]) for e in 1:2, k in axes(bx, 2)]
        check(h.ptr, ccall((:fvm_set_flux_table, LIB), Int32, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}), h.ptr, dcv, dbn))
        return nothing
    end
    model, p = s1 isa ConstantDiffusion ? (0, Float64[s.D for s in specs]) :
        s1 isa PowerDiffusion ? (2, collect(Float64, Iterators.flatten((s.D0, s.m, Float64(s.use_abs)) for s in specs))) :
        s1 isa AdvectionDiffusionFlux ? (3, collect(Float64, Iterators.flatten((s.D, s.nu_x, s.nu_y) for s in specs))) :
        s1 isa KellerSegelFlux ? (4, [s1.c, s1.D]) : unsupported("flux / diffusion", s1)
    check(h.ptr, ccall((:fvm_set_flux, LIB), Int32, (Ptr{Cvoid}, Int32, Ptr{Float64}, Int32), h.ptr, model, p, length(p)))
    return nothing
end

function set_source!(h::Handle, tri::Triangulation, specs)
    all(s -> s isa ZeroSource, specs) && return nothing
    # the reference's default source is an anonymous closure returning zero(eltype(u)) (src/problem.jl:123): not a registry
    # spec, so it is rejected like any other closure -- pass ZeroSource() explicitly
    for s in specs
        s isa DeviceFunctor || unsupported("source", s)
    end
    s1 = specs[1]
    if any(s -> s isa TabulatedSource, specs)
        all(s -> s isa Union{TabulatedSource, ZeroSource}, specs) || throw(ArgumentError("tabulated sources can only be mixed with ZeroSource"))
        tab = zeros(Float64, h.neq, h.n)                                                # column-major (neq, N) == C [N][neq]
        for (v, s) in enumerate(specs), i in 1:h.n
            s isa TabulatedSource || continue
            x, y = getxy(get_point(tri, i))
            tab[v, i] = s.fn(x, y)
        end
        check(h.ptr, ccall((:fvm_set_source_table, LIB), Int32, (Ptr{Cvoid}, Ptr{Float64}), h.ptr, tab))
        return nothing
    end
    model, p = if s1 isa GrayScottSource
        (4, [s1.b, s1.d])
    elseif s1 isa BrusselatorSource
        (5, Float64[])
    elseif s1 isa KellerSegelSource
        (6, [s1.a])
    elseif all(s -> s isa Union{LinearSource, ZeroSource}, specs)
        (1, collect(Float64, Iterators.flatten(s isa LinearSource ? (s.lam, s.mu) : (0.0, 0.0) for s in specs)))
    elseif all(s -> s isa Union{LogisticSource, ZeroSource}, specs)
        (2, Float64[s isa LogisticSource ? s.lam : 0.0 for s in specs])
    else
        throw(ArgumentError("this combination of per-species source models is not compiled"))
    end
    check(h.ptr, ccall((:fvm_set_source, LIB), Int32, (Ptr{Cvoid}, Int32, Ptr{Float64}, Int32), h.ptr, model, p, length(p)))
    return nothing
end

# ---------------------------------------------------------------------------------------------------------
# FVMProblem / FVMSystem: the params object and the RHS
# ---------------------------------------------------------------------------------------------------------
subproblems(prob::FVMProblem) = (prob,)
subproblems(prob::FVMSystem) = prob.problems                                           # src/problem.jl:233-279
subproblems(prob::SteadyFVMProblem) = subproblems(prob.problem)

"""
    get_cuda_parameters(prob; tile_triangles = 0, geometry_mode = 1, device = 0)

Sibling of `get_multithreading_parameters` (src/solve.jl:1-27): flattens `prob` (an `FVMProblem`, an `FVMSystem` or a
`SteadyFVMProblem` around either) into a device handle once and returns the `p` of `fvm_eqs!(du, u, p, t)`.
"""
function get_cuda_parameters(prob::Union{FVMProblem, FVMSystem, SteadyFVMProblem}; tile_triangles = 0, geometry_mode = 1, device = 0)
    probs = subproblems(prob)
    mesh = probs[1].mesh
    h = Handle(mesh.triangulation, length(probs); device)
    for (v, p) in enumerate(probs)
        set_conditions!(h, v - 1, p.conditions)
    end
    set_flux!(h, mesh, map(p -> flux_spec(p.flux_function), collect(probs)))
    set_source!(h, mesh.triangulation, map(p -> p.source_function, collect(probs)))
    check(h.ptr, ccall((:fvm_finalize, LIB), Int32, (Ptr{Cvoid}, Int32, Int32), h.ptr, tile_triangles, geometry_mode))
    return (prob = prob, parallel = Val(:cuda), handle = h)
end

const CudaParams = NamedTuple{(:prob, :parallel, :handle)}

"`fvm_eqs!` for `p.parallel == Val(:cuda)` (src/equations/main_equations.jl:28-35).  `u`: Vector (scalar) or Matrix(neq, N)."
function FVM.fvm_eqs!(du::Array{Float64}, u::Array{Float64}, p::CudaParams, t)
    check(p.handle.ptr, ccall((:fvm_rhs, LIB), Int32, (Ptr{Cvoid}, Float64, Ptr{Float64}, Ptr{Float64}, Int32),
        p.handle.ptr, Float64(t), u, du, 0))
    return du
end

"Dirichlet callback body (src/equations/dirichlet.jl:78-86) for integrators whose `p` is a `CudaParams`."
function FVM.update_dirichlet_nodes!(integrator, p::CudaParams = integrator.p)
    check(p.handle.ptr, ccall((:fvm_apply_dirichlet, LIB), Int32, (Ptr{Cvoid}, Float64, Ptr{Float64}, Int32),
        p.handle.ptr, Float64(integrator.t), integrator.u, 0))
    return nothing
end

"Page-locks the arrays an integrator passes to `fvm_eqs!` / `mul!` (full PCIe rate, pipelined copies); undo with `unpin!`."
pin!(a::Array{Float64}) = (ccall((:fvm_host_register, LIB), Int32, (Ptr{Cvoid}, Int64), a, sizeof(a)) == 0 || error("fvm_host_register failed"); a)
unpin!(a::Array{Float64}) = (ccall((:fvm_host_unregister, LIB), Int32, (Ptr{Cvoid},), a); a)

csr_to_csc(n, rowptr, col, val) = sparse(reduce(vcat, [fill(i, rowptr[i + 1] - rowptr[i]) for i in 1:n]; init = Int[]), Int.(col) .+ 1, val, n, n)

"`jac_prototype` of the ODEFunction (jacobian_sparsity, src/solve.jl:50-131) straight from the device pattern."
function cuda_jac_prototype(p::CudaParams)
    h = p.handle
    n = Ref{Int64}(0); nnz = Ref{Int64}(0)
    check(h.ptr, ccall((:fvm_get_jacobian_size, LIB), Int32, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}), h.ptr, n, nnz))
    rowptr = Vector{Int32}(undef, n[] + 1); col = Vector{Int32}(undef, nnz[])
    check(h.ptr, ccall((:fvm_get_jacobian_csr, LIB), Int32, (Ptr{Cvoid}, Ptr{Int32}, Ptr{Int32}, Ptr{Float64}), h.ptr, rowptr, col, C_NULL))
    return csr_to_csc(n[], rowptr, col, ones(Float64, nnz[]))
end

"Analytic Jacobian d(du)/du at (u, t) on the device; returns a SparseMatrixCSC on the `cuda_jac_prototype` pattern."
function fvm_jacobian(u::Array{Float64}, p::CudaParams, t)
    h = p.handle
    check(h.ptr, ccall((:fvm_jacobian, LIB), Int32, (Ptr{Cvoid}, Float64, Ptr{Float64}, Int32), h.ptr, Float64(t), u, 0))
    n = Ref{Int64}(0); nnz = Ref{Int64}(0)
    check(h.ptr, ccall((:fvm_get_jacobian_size, LIB), Int32, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}), h.ptr, n, nnz))
    rowptr = Vector{Int32}(undef, n[] + 1); col = Vector{Int32}(undef, nnz[]); val = Vector{Float64}(undef, nnz[])
    check(h.ptr, ccall((:fvm_get_jacobian_csr, LIB), Int32, (Ptr{Cvoid}, Ptr{Int32}, Ptr{Int32}, Ptr{Float64}), h.ptr, rowptr, col, val))
    return csr_to_csc(n[], rowptr, col, val)
end

"Device-resident `solve(prob, Tsit5(); adaptive = false, dt, saveat)` on `fvm_eqs!` (use_operator = false) or on a template."
function tsit5!(u::Array{Float64}, h::Handle, t0, t1, dt; use_operator = false, saveat = Float64[])
    ts = collect(Float64, saveat)
    us = Matrix{Float64}(undef, length(u), length(ts))
    check(h.ptr, ccall((:fvm_tsit5, LIB), Int32,
        (Ptr{Cvoid}, Int32, Ptr{Float64}, Float64, Float64, Float64, Int64, Ptr{Float64}, Ptr{Float64}, Int32),
        h.ptr, use_operator, u, t0, t1, dt, length(ts), ts, us, 0))
    return u, us
end

"Device-resident adaptive `solve(prob, Tsit5(); abstol, reltol, saveat)`; returns (u, us, n_accept, n_reject)."
function tsit5_adaptive!(u::Array{Float64}, h::Handle, t0, t1; abstol = 1e-6, reltol = 1e-3, dt0 = 0.0, use_operator = false, saveat = Float64[])
    ts = collect(Float64, saveat)
    us = Matrix{Float64}(undef, length(u), length(ts))
    na = Ref{Int64}(0); nr = Ref{Int64}(0)
    check(h.ptr, ccall((:fvm_tsit5_adaptive, LIB), Int32,
        (Ptr{Cvoid}, Int32, Ptr{Float64}, Float64, Float64, Float64, Float64, Float64, Int64, Ptr{Float64}, Ptr{Float64}, Int32, Ptr{Int64}, Ptr{Int64}),
        h.ptr, use_operator, u, t0, t1, abstol, reltol, dt0, length(ts), ts, us, 0, na, nr))
    return u, us, na[], nr[]
end

# ---------------------------------------------------------------------------------------------------------
# templates (src/specific_problems): assembly on the device, the operator, the Krylov solve
# ---------------------------------------------------------------------------------------------------------
template_id(::DiffusionEquation) = 0
template_id(::LinearReactionDiffusionEquation) = 1
template_id(::MeanExitTimeProblem) = 2
template_id(::PoissonsEquation) = 3
template_id(::LaplacesEquation) = 4

"""
    FVMCudaOperator(prob::AbstractFVMTemplate; reference_quirks = true)

Assembles the template's `A` and `b` on the device (abstract_templates.jl:73-325) and stands in for
`MatrixOperator(sparse(Afull))` (diffusion_equation.jl:82-94): `mul!(du, A, u)` acts on the reference's augmented
state `ũ = [u; 1]`, `Ã = [A b; 0 0]`.
"""
struct FVMCudaOperator
    handle::Handle
    n::Int
    b::Vector{Float64}
end
Base.size(A::FVMCudaOperator) = (A.n + 1, A.n + 1)
Base.size(A::FVMCudaOperator, i::Integer) = i <= 2 ? A.n + 1 : 1
Base.eltype(::FVMCudaOperator) = Float64

function FVMCudaOperator(prob::FVM.AbstractFVMTemplate; reference_quirks = true, source_function = nothing,
        source_parameters = nothing, tile_triangles = 0, device = 0)
    mesh, conds = prob.mesh, prob.conditions
    tri = mesh.triangulation
    h = Handle(tri, 1; device)
    set_conditions!(h, 0, conds; register_functions = false)       # templates evaluate their condition functions on the host
    check(h.ptr, ccall((:fvm_finalize, LIB), Int32, (Ptr{Cvoid}, Int32, Int32), h.ptr, tile_triangles == 0 ? 4096 : tile_triangles, 1))
    D, Dp = prob.diffusion_function, prob.diffusion_parameters
    xs, ys = cv_edge_midpoints(mesh, h); bx, by = boundary_quarter_points(tri, h)
    dcv = Float64[D(xs[e, k], ys[e, k], Dp) for e in 1:3, k in axes(xs, 2)]
    dbn = Float64[D(bx[e, k], by[e, k], Dp) for e in 1:2, k in axes(bx, 2)]
    node_value = zeros(Float64, h.n)
    if !(prob isa MeanExitTimeProblem)                              # BC functions are never evaluated there (mean_exit_time.jl:72-74)
        for dict in (conds.dudt_nodes, conds.dirichlet_nodes), (i, fidx) in dict
            x, y = getxy(get_point(tri, i))
            node_value[i] = FVM.eval_condition_fnc(conds, fidx, x, y, nothing, nothing)   # t = u = nothing (abstract_templates.jl:112,130)
        end
    end
    edge_value = zeros(Float64, 2, length(h.edges))
    if !(prob isa MeanExitTimeProblem)
        for (k, e) in enumerate(h.edges)
            haskey(conds.neumann_edges, e) || continue
            fidx = conds.neumann_edges[e]
            edge_value[1, k] = FVM.eval_condition_fnc(conds, fidx, bx[1, k], by[1, k], nothing, nothing)
            edge_value[2, k] = FVM.eval_condition_fnc(conds, fidx, bx[2, k], by[2, k], nothing, nothing)
        end
    end
    source = C_NULL
    if source_function !== nothing                                  # Poisson f(x), linear reaction-diffusion diagonal term
        source = Float64[(xy = getxy(get_point(tri, i)); source_function(xy[1], xy[2], source_parameters)) for i in 1:h.n]
    end
    check(h.ptr, ccall((:fvm_assemble, LIB), Int32,
        (Ptr{Cvoid}, Int32, Float64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Int32),
        h.ptr, template_id(prob), 1.0, dcv, dbn, node_value, edge_value, source, reference_quirks ? 1 : 0))
    _, b = get_csr(h)
    return FVMCudaOperator(h, h.n, b)
end

"A and b of an assembled template in the caller's numbering (`prob.A`, `prob.b`; e.g. diffusion_equation.jl:96-100)."
function get_csr(h::Handle)
    n = Ref{Int64}(0); nnz = Ref{Int64}(0)
    check(h.ptr, ccall((:fvm_get_csr_size, LIB), Int32, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}), h.ptr, n, nnz))
    rowptr = Vector{Int32}(undef, n[] + 1); col = Vector{Int32}(undef, nnz[]); val = Vector{Float64}(undef, nnz[]); b = Vector{Float64}(undef, n[])
    check(h.ptr, ccall((:fvm_get_csr, LIB), Int32, (Ptr{Cvoid}, Ptr{Int32}, Ptr{Int32}, Ptr{Float64}, Ptr{Float64}), h.ptr, rowptr, col, val, b))
    return csr_to_csc(n[], rowptr, col, val), b
end

"`mul!(du, A, u)` of the MatrixOperator (diffusion_equation.jl:93-94) on the augmented state: du[1:n] = A u[1:n] + u[n+1] b, du[n+1] = 0."
function LinearAlgebra.mul!(du::Vector{Float64}, A::FVMCudaOperator, u::Vector{Float64})
    length(u) == A.n + 1 == length(du) || throw(DimensionMismatch("the template state carries a trailing 1 (diffusion_equation.jl:82)"))
    last = u[end]
    check(A.handle.ptr, ccall((:fvm_spmv, LIB), Int32, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Int32, Int32),
        A.handle.ptr, u, du, last == 1.0 ? 1 : 0, 0))
    if last != 1.0 && last != 0.0
        @views du[1:A.n] .+= last .* A.b
    end
    du[end] = 0.0
    return du
end
Base.:*(A::FVMCudaOperator, u::Vector{Float64}) = mul!(similar(u), A, u)

"Jacobi-preconditioned Krylov solve of a steady template on the device, in place of `KLUFactorization()`."
struct FVMCudaKrylov
    method::Symbol      # :pcg (symmetric operators: constant D, no Constrained edges) or :bicgstab
    rtol::Float64
    maxiter::Int
end
FVMCudaKrylov(; method = :bicgstab, rtol = 1e-10, maxiter = 100_000) = FVMCudaKrylov(method, rtol, maxiter)

"`solve(prob::AbstractFVMTemplate, alg)` (abstract_templates.jl:58-60) for the steady templates."
function CommonSolve.solve(prob::Union{PoissonsEquation, LaplacesEquation, MeanExitTimeProblem}, alg::FVMCudaKrylov;
        source_function = nothing, source_parameters = nothing, x0 = nothing, kwargs...)
    A = FVMCudaOperator(prob; source_function, source_parameters, kwargs...)
    x = x0 === nothing ? zeros(Float64, A.n) : collect(Float64, x0)
    iters = Ref{Int32}(0); relres = Ref{Float64}(0.0)
    check(A.handle.ptr, ccall((:fvm_krylov, LIB), Int32, (Ptr{Cvoid}, Int32, Ptr{Float64}, Float64, Int32, Ptr{Int32}, Ptr{Float64}, Int32),
        A.handle.ptr, alg.method === :pcg ? 0 : 1, x, alg.rtol, alg.maxiter, iters, relres, 0))
    return (u = x, iters = Int(iters[]), resid = relres[], retcode = relres[] <= 10 * alg.rtol ? :Success : :MaxIters)
end

# ---------------------------------------------------------------------------------------------------------
# solve(SteadyFVMProblem(prob), alg) on the device                                        src/solve.jl:209-220
# ---------------------------------------------------------------------------------------------------------
"Newton-Raphson with the Newton systems solved by Jacobi-preconditioned BiCGStab on the device Jacobian (fvm_newton)."
Base.@kwdef struct FVMCudaNewton
    abstol::Float64 = 1e-11
    reltol::Float64 = 1e-11
    maxiters::Int = 50
    lin_rtol::Float64 = 1e-13
    lin_maxiters::Int = 20_000
end

function CommonSolve.solve(prob::SteadyFVMProblem, alg::FVMCudaNewton; p::CudaParams = get_cuda_parameters(prob.problem), kwargs...)
    h = p.handle
    u = collect(Float64, prob.problem.initial_condition)
    iters = Ref{Int32}(0); lin = Ref{Int64}(0); res = Ref{Float64}(0.0); res0 = Ref{Float64}(0.0)
    check(h.ptr, ccall((:fvm_newton, LIB), Int32,
        (Ptr{Cvoid}, Float64, Ptr{Float64}, Float64, Float64, Int32, Float64, Int32, Ptr{Int32}, Ptr{Float64}, Ptr{Float64}, Ptr{Int64}, Int32),
        h.ptr, Float64(prob.problem.initial_time), u, alg.abstol, alg.reltol, Int32(alg.maxiters), alg.lin_rtol, Int32(alg.lin_maxiters),
        iters, res, res0, lin, 0))
    ok = res[] <= alg.abstol + alg.reltol * res0[]
    return (u = u, iters = Int(iters[]), linear_iters = Int(lin[]), resid = res[], retcode = ok ? :Success : :MaxIters)
end

# ---------------------------------------------------------------------------------------------------------
# post-processing on the device (pl_interpolate src/utils.jl:23-27, compute_flux src/problem.jl:458-487)
# ---------------------------------------------------------------------------------------------------------
"u (or q . n when `normals` is given) at points lying in the given triangles (indices into the handle's triangle order)."
function eval_points(p::CudaParams, u::Array{Float64}, t, tri_idx::Vector{Int32}, xy::Matrix{Float64}; normals = nothing)
    n = length(tri_idx)
    out = Matrix{Float64}(undef, p.handle.neq, n)
    check(p.handle.ptr, ccall((:fvm_eval_points, LIB), Int32,
        (Ptr{Cvoid}, Float64, Ptr{Float64}, Int32, Int64, Ptr{Int32}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
        p.handle.ptr, Float64(t), u, 0, n, tri_idx, xy, normals === nothing ? C_NULL : normals, out))
    return out
end

"Geometry read-back for parity checks against `mesh.cv_volumes` / `mesh.triangle_props` (src/geometry.jl:21-49)."
function get_geometry(h::Handle)
    T = length(h.triangles)
    V = Vector{Float64}(undef, h.n); s = Matrix{Float64}(undef, 9, T); mid = Array{Float64}(undef, 2, 3, T)
    nrm = Array{Float64}(undef, 2, 3, T); len = Matrix{Float64}(undef, 3, T)
    check(h.ptr, ccall((:fvm_get_geometry, LIB), Int32, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
        h.ptr, V, s, mid, nrm, len))
    return V, s, mid, nrm, len
end

# ---------------------------------------------------------------------------------------------------------
# FVMWIRE containers: the importer / exporter for DelaunayTriangulation objects and solutions
# ---------------------------------------------------------------------------------------------------------
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
    pts = reshape(collect(Float64, Iterators.flatten(DT.each_point(tri))), 2, :)
    T = reshape(collect(Int32, Iterators.flatten(triangle_vertices(t) for t in each_solid_triangle(tri))), 3, :)
    edges = collect(keys(get_boundary_edge_map(tri)))
    wire_put(w, "points", pts); wire_put(w, "triangles", T); wire_put(w, "index_base", Int32[1])
    wire_put(w, "boundary_edges", reshape(collect(Int32, Iterators.flatten(edges)), 2, :))
    # section of an edge (u, v): the ghost vertex on its other side is -section (src/conditions.jl:507-515)
    wire_put(w, "boundary_edge_section", Int32[-get_adjacent(tri, v, u) - 1 for (u, v) in edges])
    wire_put(w, "num_sections", Int32[length(DT.get_ghost_vertex_map(tri))])
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

"Raw handle straight from a mesh container (no Triangulation object on the Julia side); continue with the setters."
function handle_from_wire(path::String, neq::Integer; device = 0)
    out = Ref{Ptr{Cvoid}}(C_NULL)
    rc = ccall((:fvm_create_from_wire, LIB), Int32, (Cstring, Int32, Int32, Ptr{Ptr{Cvoid}}), path, neq, device, out)
    rc == 0 || throw(FVMCudaError(rc, unsafe_string(ccall((:fvm_wire_last_error, LIB), Cstring, (Ptr{Cvoid},), C_NULL))))
    return out[]
end

end # module
