# LQRB200.jl — `ccall`-only Julia shim over liblqrb200.so (include/lqrb200.h).
#
# STATUS: static artefact.  Julia is not installed in the build image or on the GPU box, so this file
# has never been executed; the executable tests drive the same C ABI from Python (ctypes).  It exists to
# show, name for name, how LQR.jl's hot path is swapped for the B200 library:
#
#   reference (bjack205/LQR.jl)                               this shim
#   LQRProblem, size, num_vars   src/lqr_problem.jl:1-25       BatchedLQRProblem, Base.size, num_vars
#   DPSolver, solve!             src/dynamic_programming.jl    DPSolver, solve!
#   BlockCholesky, cholesky!,    src/block_cholesky.jl:19-101  BlockCholesky, cholesky!, ldiv!, \
#     ldiv!, \
#   CholeskySolver._solve!       src/cholesky_solver.jl:166    _solve!(::BatchedCholeskySolver)
#   second_order_correction!     src/cholesky_solver.jl:254    second_order_correction!
#
# Arrays are ordinary Julia `Array{Float64}` whose LAST axis is the batch: A is n×n×(N-1)×batch.  That is
# exactly the ABI's "instance-major" layout, so no copies are made on the Julia side; the library moves
# the data to the GPU, repacks it batch-minor, solves, and writes the results back.
module LQRB200

using LinearAlgebra

const lib = get(ENV, "LQRB200_LIB", joinpath(@__DIR__, "..", "liblqrb200.so"))

const HESS_DENSE, HESS_BLOCKDIAG, HESS_DIAG = Int32(0), Int32(1), Int32(2)
const FLAG_SOC, FLAG_LTI = Int32(1), Int32(2)

struct LQRBError <: Exception
    code::Int32
    msg::String
end

mutable struct Handle
    ptr::Ptr{Cvoid}
    function Handle(device::Integer=0)
        ref = Ref{Ptr{Cvoid}}(C_NULL)
        rc = ccall((:lqrb_create, lib), Int32, (Ref{Ptr{Cvoid}}, Int32), ref, device)
        rc == 0 || throw(LQRBError(rc, "lqrb_create failed (an sm_100 GPU is required; there is no CPU fallback)"))
        h = new(ref[])
        finalizer(h -> ccall((:lqrb_destroy, lib), Int32, (Ptr{Cvoid},), h.ptr), h)
        return h
    end
end

function check(h::Handle, rc::Int32)
    rc == 0 && return nothing
    msg = unsafe_string(ccall((:lqrb_last_error_string, lib), Cstring, (Ptr{Cvoid},), h.ptr))
    throw(LQRBError(rc, msg))
end

ptr_or_null(a::Nothing) = Ptr{Float64}(C_NULL)
ptr_or_null(a::Array{Float64}) = pointer(a)

# ------------------------------------------------------------------ LQRProblem (src/lqr_problem.jl:1-25)
struct BatchedLQRProblem
    Qf::Array{Float64,3}   # n×n×batch
    Q::Array{Float64}      # n×n×batch (LTI) or n×n×(N-1)×batch (LTV)
    R::Array{Float64}
    A::Array{Float64}
    B::Array{Float64}
    x0::Matrix{Float64}    # n×batch
    q::Union{Nothing,Array{Float64}}
    r::Union{Nothing,Array{Float64}}
    qf::Union{Nothing,Matrix{Float64}}
    tf::Float64
    N::Int
end
islti(p::BatchedLQRProblem) = ndims(p.A) == 3
Base.size(p::BatchedLQRProblem) = (size(p.A, 1), size(p.B, 2), p.N)
batchsize(p::BatchedLQRProblem) = size(p.x0, 2)
num_vars(p) = ((n, m, N) = size(p); N * n + (N - 1) * m)

# ------------------------------------------------------------------ DPSolver (src/dynamic_programming.jl)
struct DPSolver
    handle::Handle
end
DPSolver(prob::BatchedLQRProblem; device=0) = DPSolver(Handle(device))

struct LQRSolution
    Z::Matrix{Float64}           # NN×batch, Primals order [x1;u1;…;xN] (src/lqr_problem.jl:46-73)
    K::Array{Float64,4}          # m×n×(N-1)×batch
    d::Array{Float64,3}          # m×(N-1)×batch
    info::Vector{Int32}
end
function LQRSolution(prob::BatchedLQRProblem)
    n, m, N = size(prob); b = batchsize(prob)
    LQRSolution(zeros(num_vars(prob), b), zeros(m, n, N - 1, b), zeros(m, N - 1, b), zeros(Int32, b))
end

"solve!(sol, solver::DPSolver, prob): src/dynamic_programming.jl:54-72"
function solve!(sol::LQRSolution, solver::DPSolver, prob::BatchedLQRProblem)
    n, m, N = size(prob)
    flags = islti(prob) ? FLAG_LTI : Int32(0)
    GC.@preserve sol prob begin
        rc = ccall((:lqrb_riccati_f64, lib), Int32,
            (Ptr{Cvoid}, Int32, Int32, Int32, Int64, Int32,
             Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64},
             Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Int32}),
            solver.handle.ptr, n, m, N, batchsize(prob), flags,
            prob.A, prob.B, prob.Q, prob.R, ptr_or_null(prob.q), ptr_or_null(prob.r),
            prob.Qf, ptr_or_null(prob.qf), prob.x0, sol.Z, sol.K, sol.d, sol.info)
    end
    check(solver.handle, rc)
    return sol
end

# ------------------------------------------------------------------ BlockCholesky (src/block_cholesky.jl)
struct BlockCholesky
    handle::Handle
    M::Array{Float64,3}     # (n+m)×(n+m)×batch: upper factor, or reciprocals on the diagonal (diag mode)
    n::Int
    m::Int
    mode::Int32
    info::Vector{Int32}
end
function BlockCholesky(handle::Handle, n::Int, m::Int, batch::Int; diag=false, block_diag=false)
    mode = diag ? HESS_DIAG : (block_diag ? HESS_BLOCKDIAG : HESS_DENSE)
    BlockCholesky(handle, zeros(n + m, n + m, batch), n, m, mode, zeros(Int32, batch))
end

"cholesky!(chol, A, B[, C]): src/block_cholesky.jl:55-91"
function LinearAlgebra.cholesky!(chol::BlockCholesky, A::Array{Float64,3}, B::Array{Float64,3},
                                 C::Union{Nothing,Array{Float64,3}}=nothing)
    GC.@preserve chol A B C begin
        rc = ccall((:lqrb_block_cholesky_f64, lib), Int32,
            (Ptr{Cvoid}, Int32, Int32, Int64, Int32, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Int32}),
            chol.handle.ptr, chol.n, chol.m, size(chol.M, 3), chol.mode, A, B, ptr_or_null(C), chol.M, chol.info)
    end
    check(chol.handle, rc)
    return chol
end

"ldiv!(chol, b): src/block_cholesky.jl:93-96; b is (n+m)×nrhs×batch"
function LinearAlgebra.ldiv!(chol::BlockCholesky, b::Array{Float64,3})
    GC.@preserve chol b begin
        rc = ccall((:lqrb_block_ldiv_f64, lib), Int32,
            (Ptr{Cvoid}, Int32, Int32, Int64, Int32, Ptr{Float64}, Int32, Ptr{Float64}),
            chol.handle.ptr, chol.n, chol.m, size(chol.M, 3), chol.mode, chol.M, size(b, 2), b)
    end
    check(chol.handle, rc)
    return b
end
Base.:\(chol::BlockCholesky, b::Array{Float64,3}) = ldiv!(chol, copy(b))

# ------------------------------------------------------------------ CholeskySolver (src/cholesky_solver.jl)
# The reference fills Jinv / conSet.blocks through TrajOptCore (update!, :155-164); a batched caller
# hands over the same linearised blocks as plain arrays with a trailing batch axis.
mutable struct BatchedCholeskySolver
    handle::Handle
    n::Int; m::Int; N::Int
    p::Vector{Int32}             # stage-constraint rows per knot (src/conblocks.jl:74-96)
    hess_mode::Int32
    Q::Array{Float64,4}; R::Array{Float64,4}; Hux::Union{Nothing,Array{Float64,4}}
    q::Array{Float64,3}; r::Array{Float64,3}
    A::Array{Float64,4}; B::Array{Float64,4}; d::Array{Float64,3}
    D2::Union{Nothing,Matrix{Float64}}     # nothing ⇒ [-I 0]
    C::Matrix{Float64}; c::Matrix{Float64} # concatenated blocks × batch
    δZ::Matrix{Float64}; λ::Matrix{Float64}; res::Matrix{Float64}; info::Vector{Int32}
end
Base.size(s::BatchedCholeskySolver) = (s.n, s.m, s.N)

function _call_kkt!(s::BatchedCholeskySolver, flags::Int32)
    GC.@preserve s begin
        rc = ccall((:lqrb_kkt_solve_f64, lib), Int32,
            (Ptr{Cvoid}, Int32, Int32, Int32, Int64, Ptr{Int32}, Int32, Int32,
             Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64},
             Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Int32}),
            s.handle.ptr, s.n, s.m, s.N, size(s.δZ, 2), s.p, s.hess_mode, flags,
            s.Q, s.R, ptr_or_null(s.Hux), s.q, s.r, s.A, s.B, s.d, ptr_or_null(s.D2), s.C, s.c,
            s.δZ, s.λ, s.res, s.info)
    end
    check(s.handle, rc)
    return s
end

"_solve!(solver): src/cholesky_solver.jl:166-182 (Schur factors, block Cholesky, substitutions, primals)"
_solve!(s::BatchedCholeskySolver) = _call_kkt!(s, Int32(0))
"second_order_correction!: src/cholesky_solver.jl:254-273 (the Ginv=false chain)"
second_order_correction!(s::BatchedCholeskySolver) = _call_kkt!(s, FLAG_SOC)
"residual(solver; recalculate): src/cholesky_solver.jl:238-252.  recalculate=true evaluates res on the device
from the solver's current blocks and the multipliers kept from the last solve (calc_residual!, :201-236)."
function residual(s::BatchedCholeskySolver; recalculate::Bool=false)
    n, m, N = size(s)
    if recalculate
        norms = zeros(size(s.res, 2))
        GC.@preserve s norms begin
            rc = ccall((:lqrb_kkt_residual_f64, lib), Int32,
                (Ptr{Cvoid}, Int32, Int32, Int32, Int64, Ptr{Int32}, Int32,
                 Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64},
                 Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
                s.handle.ptr, s.n, s.m, s.N, size(s.res, 2), s.p, Int32(0),
                s.q, s.r, s.A, s.B, ptr_or_null(s.D2), s.C, s.λ, s.res, norms)
        end
        check(s.handle, rc)
        return norms
    end
    map(1:size(s.res, 2)) do i
        r = view(s.res, :, i)
        norm([norm(view(r, (k - 1) * (n + m) .+ (1:(k < N ? n + m : n)))) for k = 1:N])
    end
end
get_step(s::BatchedCholeskySolver) = s.δZ
get_multipliers(s::BatchedCholeskySolver) = s.λ

export Handle, BatchedLQRProblem, DPSolver, LQRSolution, solve!, BlockCholesky, BatchedCholeskySolver,
       _solve!, second_order_correction!, residual, get_step, get_multipliers, num_vars

end # module
