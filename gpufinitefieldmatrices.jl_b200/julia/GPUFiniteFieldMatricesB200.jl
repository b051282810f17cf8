# GPUFiniteFieldMatricesB200.jl -- drop-in Julia host layer for the B200-native library.
#
# Same exported names, argument meaning and error behaviour as the reference module
# (reference src/GPUFiniteFieldMatrices.jl:36-60), but every operation is one `ccall` into libgffm.so
# (include/gffm.h): no CUDA.jl kernels, no cuBLAS, no CPU fallback.  This file is declarative on purpose -- one
# `ccall` per exported function -- and is mirrored 1:1 by the ctypes binding (capi.py / cumodmatrix.py) that the
# test-suite executes, because no Julia toolchain exists in the build image.
module GPUFiniteFieldMatricesB200

using LinearAlgebra
import LinearAlgebra: mul!, rmul!, lmul!, transpose
import Base: size, length, eltype, getindex, setindex!, show, +, -, *, /, ^, copy, copy!, copyto!, fill!, Array

const libgffm = get(ENV, "GFFM_LIB", joinpath(@__DIR__, "..", "lib", "libgffm.so"))
const TILE_WIDTH = 32            # reference CuModMatrix.jl:2
const DEFAULT_TYPE = Float32     # reference CuModMatrix.jl:3

# ---- exceptions (reference CuModMatrix.jl:5-31; MatrixNotInvertibleException is used but never defined there, :485)
struct CuModArraySizeMismatchException <: Exception; message::String; end
struct CuModArrayModulusMismatchException <: Exception; message::String; end
struct CuModMatrixTooLargeException <: Exception; message::String; end
struct CuModMatrixNotSquareException <: Exception; message::String; end
struct CuModMatrixModulusNotPrimeException <: Exception; message::String; end
struct InverseOverflowError <: Exception; message::String; end
struct InverseNotDefinedException <: Exception; message::String; end
struct MatrixNotInvertibleException <: Exception; message::String; end

last_error() = unsafe_string(ccall((:gffm_last_error, libgffm), Cstring, ()))
version() = unsafe_string(ccall((:gffm_version, libgffm), Cstring, ()))

function check(st::Int32)
    st == 0 && return nothing
    msg = last_error()
    st == 2 && throw(CuModArraySizeMismatchException(msg))
    (st == 3 || st == 4) && throw(CuModArrayModulusMismatchException(msg))
    st == 5 && throw(CuModMatrixNotSquareException(msg))
    st == 6 && throw(MatrixNotInvertibleException(msg))
    st == 7 && throw(InverseNotDefinedException(msg))
    st == 11 && throw(InexactError(:convert, Integer, msg))
    st == 13 && throw(CuModMatrixModulusNotPrimeException(msg))
    st == 1 && throw(ArgumentError(msg))
    error("libgffm status $st: $msg")
end

# ---- context (CUDA.jl's task-local device/stream state in the reference) ---------------------------------------
mutable struct Context
    h::Ptr{Cvoid}
    function Context(device::Integer=0)
        r = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:gffm_create, libgffm), Int32, (Int32, Ref{Ptr{Cvoid}}), device, r))
        # No finalizer: Julia does not order finalizers, and a matrix finalizer (gffm_mat_destroy) needs its context's stream.
        # Every CuModArray holds a reference to its Context, so an unreachable Context has no live matrices; its few MB of
        # workspaces are released by close(ctx) or at process exit.
        return new(r[])
    end
end
Base.close(c::Context) = (c.h == C_NULL || ccall((:gffm_destroy, libgffm), Int32, (Ptr{Cvoid},), c.h); c.h = C_NULL; nothing)
const _ctx = Ref{Union{Nothing,Context}}(nothing)
default_context() = (_ctx[] === nothing && (_ctx[] = Context(0)); _ctx[])
device_count() = (r = Ref{Int32}(0); check(ccall((:gffm_device_count, libgffm), Int32, (Ref{Int32},), r)); Int(r[]))
synchronize(c::Context=default_context()) = check(ccall((:gffm_sync, libgffm), Int32, (Ptr{Cvoid},), c.h))
set_stream!(c::Context, s::Ptr{Cvoid}) = check(ccall((:gffm_set_stream, libgffm), Int32, (Ptr{Cvoid}, Ptr{Cvoid}), c.h, s))
get_stream(c::Context) = (r = Ref{Ptr{Cvoid}}(C_NULL); check(ccall((:gffm_get_stream, libgffm), Int32, (Ptr{Cvoid}, Ref{Ptr{Cvoid}}), c.h, r)); r[])
alloc_stats(c::Context=default_context()) = (b = Ref{Int64}(0); k = Ref{Int64}(0); check(ccall((:gffm_alloc_stats, libgffm), Int32, (Ptr{Cvoid}, Ref{Int64}, Ref{Int64}), c.h, b, k)); (b[], k[]))
launch_count(c::Context=default_context()) = (r = Ref{Int64}(0); check(ccall((:gffm_launch_count, libgffm), Int32, (Ptr{Cvoid}, Ref{Int64}), c.h, r)); r[])
set_gemm_ctas!(c::Context, n::Integer) = check(ccall((:gffm_set_gemm_ctas, libgffm), Int32, (Ptr{Cvoid}, Int32), c.h, n))
set_profiling!(c::Context, on::Bool) = check(ccall((:gffm_set_profiling, libgffm), Int32, (Ptr{Cvoid}, Int32), c.h, on ? 1 : 0))
function last_timings(c::Context=default_context())
    buf = zeros(Float64, 16); n = Ref{Int32}(0)
    check(ccall((:gffm_last_timings, libgffm), Int32, (Ptr{Cvoid}, Ptr{Float64}, Int32, Ref{Int32}), c.h, buf, 16, n))
    buf[1:n[]]
end

# ---- container: struct CuModArray{T,D} (reference CuModMatrix.jl:42-46) ---------------------------------------------
_dtype(::Type{Float32}) = Int32(0); _dtype(::Type{Float64}) = Int32(1); _dtype(::Type{Int64}) = Int32(2)
_dtype(::Type{UInt32}) = Int32(3); _dtype(::Type{Int32}) = Int32(4)

mutable struct CuModArray{T,D} <: AbstractArray{T,D}      # reference CuModMatrix.jl:42
    h::Ptr{Cvoid}      # gffm_mat*
    N::Int
    ctx::Context
    function CuModArray{T,D}(h::Ptr{Cvoid}, N::Integer, ctx::Context) where {T,D}
        m = new{T,D}(h, Int(N), ctx)
        finalizer(x -> ccall((:gffm_mat_destroy, libgffm), Int32, (Ptr{Cvoid},), x.h), m)   # no Julia callbacks inside
        return m
    end
end
const CuModMatrix{T} = CuModArray{T,2}
const CuModVector{T} = CuModArray{T,1}

function _create(::Type{T}, D::Int, rows, cols, N; ctx=default_context()) where {T}
    r = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:gffm_mat_create, libgffm), Int32, (Ptr{Cvoid}, Int64, Int64, UInt64, Int32, Ref{Ptr{Cvoid}}), ctx.h, rows, cols, N, -1, r))
    CuModArray{T,D}(r[], N, ctx)
end

# host ctor CuModMatrix(A, N; mod=true, new_size=nothing, elem_type=Float32) (reference CuModMatrix.jl:53-99,143-145)
function CuModArray{T,D}(A::AbstractArray, N::Integer; mod::Bool=true, new_size=nothing, ctx=default_context()) where {T,D}
    rows = size(A, 1); cols = D == 1 ? 1 : size(A, 2)
    if new_size !== nothing
        rows = new_size[1]; cols = D == 1 ? 1 : new_size[2]
    end
    m = _create(T, D, rows, cols, N; ctx=ctx)
    S = eltype(A) <: AbstractFloat ? (eltype(A) == Float32 ? Float32 : Float64) : Int64
    host = zeros(S, rows, cols)
    r0 = min(rows, size(A, 1)); c0 = min(cols, size(A, 2))
    host[1:r0, 1:c0] .= convert.(S, A[1:r0, 1:c0])
    check(ccall((:gffm_mat_upload, libgffm), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Int32, Int64, Int32), m.h, host, _dtype(S), rows, mod ? 1 : 0))
    return m
end
CuModMatrix(A::AbstractMatrix, N::Integer; elem_type::Type=DEFAULT_TYPE, kw...) = CuModArray{elem_type,2}(A, N; kw...)
CuModVector(A::AbstractVector, N::Integer; elem_type::Type=DEFAULT_TYPE, kw...) = CuModArray{elem_type,1}(A, N; kw...)
# device-wrapper ctor (reference CuModMatrix.jl:113-121): adopt an existing UInt32 device buffer
function wrap_device(::Type{T}, ptr::Ptr{Cvoid}, rows, cols, ld, N; ctx=default_context()) where {T}
    r = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:gffm_mat_wrap, libgffm), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Int64, Int64, Int64, UInt64, Ref{Ptr{Cvoid}}), ctx.h, ptr, rows, cols, ld, N, r))
    CuModArray{T,2}(r[], N, ctx)
end

_i64(f, A) = (r = Ref{Int64}(0); check(f(A.h, r)); Int(r[]))
rows(A::CuModArray) = _i64((h, r) -> ccall((:gffm_mat_rows, libgffm), Int32, (Ptr{Cvoid}, Ref{Int64}), h, r), A)
cols(A::CuModArray) = _i64((h, r) -> ccall((:gffm_mat_cols, libgffm), Int32, (Ptr{Cvoid}, Ref{Int64}), h, r), A)
leading_dim(A::CuModArray) = _i64((h, r) -> ccall((:gffm_mat_ld, libgffm), Int32, (Ptr{Cvoid}, Ref{Int64}), h, r), A)
padding(A::CuModArray) = (r = Ref{Int32}(0); check(ccall((:gffm_mat_pad, libgffm), Int32, (Ptr{Cvoid}, Ref{Int32}), A.h, r)); Int(r[]))
modulus(A::CuModArray) = (r = Ref{UInt64}(0); check(ccall((:gffm_mat_modulus, libgffm), Int32, (Ptr{Cvoid}, Ref{UInt64}), A.h, r)); Int(r[]))
device_ptr(A::CuModArray) = (r = Ref{Ptr{Cvoid}}(C_NULL); check(ccall((:gffm_mat_device_ptr, libgffm), Int32, (Ptr{Cvoid}, Ref{Ptr{Cvoid}}), A.h, r)); r[])
size(A::CuModArray{T,2}) where {T} = (rows(A), cols(A))
size(A::CuModArray{T,1}) where {T} = (rows(A),)
size(A::CuModArray, d::Integer) = d == 1 ? rows(A) : (d == 2 ? cols(A) : 1)
length(A::CuModArray) = rows(A) * cols(A)
eltype(::CuModArray{T}) where {T} = T

# Array(A) / unsafe_Array(A) (reference CuModMatrix.jl:251-261)
function _download(A::CuModArray{T}, padded::Bool) where {T}
    r = rows(A) + (padded ? TILE_WIDTH : 0); c = cols(A) + (padded ? TILE_WIDTH : 0)
    out = zeros(T, r, c)
    check(ccall((:gffm_mat_download, libgffm), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Int32, Int64, Int32), A.h, out, _dtype(T), max(r, 1), padded ? 1 : 0))
    return out
end
Array(A::CuModArray{T,2}) where {T} = _download(A, false)
Array(A::CuModArray{T,1}) where {T} = vec(_download(A, false))
unsafe_Array(A::CuModArray) = _download(A, true)
function getindex(A::CuModArray{T}, i::Integer, j::Integer=1) where {T}
    r = Ref{Int64}(0); check(ccall((:gffm_mat_get_elem, libgffm), Int32, (Ptr{Cvoid}, Int64, Int64, Ref{Int64}), A.h, i - 1, j - 1, r)); T(r[])
end
setindex!(A::CuModArray, v, i::Integer, j::Integer=1) = check(ccall((:gffm_mat_set_elem, libgffm), Int32, (Ptr{Cvoid}, Int64, Int64, Int64), A.h, i - 1, j - 1, Int64(v)))
show(io::IO, A::CuModArray{T}) where {T} = print(io, "$(rows(A))x$(cols(A)) CuModMatrix{$T} modulo $(A.N)")
show(io::IO, ::MIME"text/plain", A::CuModArray) = show(io, A)      # AbstractArray's default display would fetch element by element
Base.IndexStyle(::Type{<:CuModArray}) = IndexCartesian()

# zeros / eye / rand (reference CuModMatrix.jl:510-556)
zeros(::Type{T}, r::Integer, c::Integer, N::Integer) where {T} = _create(T, 2, r, c, N)
eye(::Type{T}, n::Integer, N::Integer) where {T} = (m = _create(T, 2, n, n, N); check(ccall((:gffm_mat_eye, libgffm), Int32, (Ptr{Cvoid},), m.h)); m)
rand(::Type{T}, r::Integer, c::Integer, N::Integer; seed::Integer=0) where {T} =
    (m = _create(T, 2, r, c, N); check(ccall((:gffm_mat_rand, libgffm), Int32, (Ptr{Cvoid}, UInt64), m.h, seed)); m)
synth(::Type{T}, r::Integer, c::Integer, N::Integer, seed::Integer) where {T} =
    (m = _create(T, 2, r, c, N); check(ccall((:gffm_mat_synth, libgffm), Int32, (Ptr{Cvoid}, UInt64), m.h, seed)); m)
copy!(dst::CuModArray, src::CuModArray) = (check(ccall((:gffm_mat_copy, libgffm), Int32, (Ptr{Cvoid}, Ptr{Cvoid}), dst.h, src.h)); dst)
copyto!(dst::CuModArray, src::CuModArray) = copy!(dst, src)
copy(A::CuModArray{T,D}) where {T,D} = copy!(_create(T, D, rows(A), cols(A), A.N; ctx=A.ctx), A)
copy_block!(dst::CuModArray, dr, dc, src::CuModArray, sr, sc, nr, nc) =
    check(ccall((:gffm_mat_copy_block, libgffm), Int32, (Ptr{Cvoid}, Int64, Int64, Ptr{Cvoid}, Int64, Int64, Int64, Int64), dst.h, dr - 1, dc - 1, src.h, sr - 1, sc - 1, nr, nc))
fill!(A::CuModArray, v) = (check(ccall((:gffm_mat_fill, libgffm), Int32, (Ptr{Cvoid}, Int64), A.h, Int64(v))); A)
zero!(A::CuModArray) = (check(ccall((:gffm_mat_zero, libgffm), Int32, (Ptr{Cvoid},), A.h)); A)
function change_modulus_no_alloc!(A::CuModArray, N::Integer)   # reference CuModMatrix.jl:745-760
    check(ccall((:gffm_mat_set_modulus, libgffm), Int32, (Ptr{Cvoid}, UInt64, Int32), A.h, N, 1)); A.N = N; A
end
change_modulus(A::CuModArray, N::Integer) = change_modulus_no_alloc!(copy(A), N)   # :726-740
transpose(A::CuModArray{T,2}) where {T} = (B = _create(T, 2, cols(A), rows(A), A.N; ctx=A.ctx); check(ccall((:gffm_mat_transpose, libgffm), Int32, (Ptr{Cvoid}, Ptr{Cvoid}), B.h, A.h)); B)
isequal_device(A::CuModArray, B::CuModArray) = (r = Ref{Int32}(0); check(ccall((:gffm_mat_equal, libgffm), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Ref{Int32}), A.h, B.h, r)); r[] != 0)
touch!(A::CuModArray) = (check(ccall((:gffm_mat_touch, libgffm), Int32, (Ptr{Cvoid},), A.h)); A)
drop_cache!(A::CuModArray) = (check(ccall((:gffm_mat_drop_cache, libgffm), Int32, (Ptr{Cvoid},), A.h)); A)
checksum(A::CuModArray) = (r = Ref{UInt64}(0); check(ccall((:gffm_mat_checksum, libgffm), Int32, (Ptr{Cvoid}, Ref{UInt64}), A.h, r)); r[])

# ---- elementwise (reference kernel_ops/*.jl) ----------------------------------------------------------------------
_ew(op, C, A, B, s, modN) = (check(ccall((:gffm_ewise, libgffm), Int32, (Int32, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Int64, UInt64),
                                         op, C.h, A.h, B === nothing ? C_NULL : B.h, Int64(s), modN)); C)
mod_elements!(A::CuModArray, mod_N::Integer=0) = _ew(0, A, A, nothing, 0, mod_N)
add!(C, A, B, mod_N::Integer=0) = _ew(1, C, A, B, 0, mod_N)
sub!(C, A, B, mod_N::Integer=0) = _ew(2, C, A, B, 0, mod_N)
elementwise_multiply!(C, A, B, mod_N::Integer=0) = _ew(3, C, A, B, 0, mod_N)
scalar_add!(C, A, s::Number, mod_N::Integer=0) = _ew(4, C, A, nothing, s, mod_N)
scalar_sub!(C, A, s::Number, mod_N::Integer=0) = _ew(5, C, A, nothing, s, mod_N)
rscalar_sub!(C, A, s::Number, mod_N::Integer=0) = _ew(6, C, A, nothing, s, mod_N)
negate!(C, A, mod_N::Integer=0) = _ew(6, C, A, nothing, 0, mod_N)
mul!(C::CuModArray, A::CuModArray, s::Number, mod_N::Integer=0) = _ew(7, C, A, nothing, s, mod_N)
div!(C, A, s::Number, mod_N::Integer=0) = _ew(8, C, A, nothing, s, mod_N)
rmul!(A::CuModArray, s::Number) = mul!(A, A, s)
lmul!(s::Number, A::CuModArray) = mul!(A, A, s)
_like(A::CuModArray{T,D}) where {T,D} = _create(T, D, rows(A), cols(A), A.N; ctx=A.ctx)
+(A::CuModArray, B::CuModArray) = add!(_like(A), A, B)
-(A::CuModArray, B::CuModArray) = sub!(_like(A), A, B)
+(A::CuModArray, s::Number) = scalar_add!(_like(A), A, s); +(s::Number, A::CuModArray) = A + s
-(A::CuModArray, s::Number) = scalar_sub!(_like(A), A, s); -(s::Number, A::CuModArray) = rscalar_sub!(_like(A), A, s)
-(A::CuModArray) = negate!(_like(A), A)
*(A::CuModArray, s::Number) = mul!(_like(A), A, s); *(s::Number, A::CuModArray) = A * s
/(A::CuModArray, s::Number) = div!(_like(A), A, s)

# ---- modular GEMM / GEMV (reference CuModMatrix.jl:767-836, kernel_mul/stripe_mul.jl:82-244) ------------------------------
# matrix form: M (stripe width) is meaningless here and ignored; N = modulus override.  vector form: R = input bound, P = modulus.
function mul!(C::CuModArray{T,2}, A::CuModArray{T,2}, B::CuModArray{T,2}; M=nothing, N::Integer=0, mode::Integer=0, algo::Integer=0) where {T}
    check(ccall((:gffm_gemm, libgffm), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, UInt64, UInt64, Int32, Int32), C.h, A.h, B.h, 0, N, mode, algo)); C
end
mulN!(C, A, B, N::Integer) = mul!(C, A, B; N=N)
stripe_mul!(C, A, B; kw...) = mul!(C, A, B; kw...)
function mul!(z::CuModArray{T,1}, A::CuModArray{T,2}, x::CuModArray{T,1}; R::Integer=0, P::Integer=0, maxopsOverride=false) where {T}
    check(ccall((:gffm_gemv, libgffm), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, UInt64, UInt64), z.h, A.h, x.h, R, P)); z
end
# host-to-host pipelined product (uint32 residues): one call instead of CuModMatrix(A); CuModMatrix(B); mul!; Array
function mul_host!(C::Matrix{UInt32}, A::Matrix{UInt32}, B::Matrix{UInt32}, N::Integer; ctx=default_context())
    m, k = size(A); n = size(B, 2)
    check(ccall((:gffm_gemm_host, libgffm), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Int64, Ptr{Cvoid}, Int64, Ptr{Cvoid}, Int64, Int64, Int64, Int64, Int32, UInt64),
                ctx.h, C, size(C, 1), A, m, B, k, m, n, k, 3, N)); C
end
# multi-GPU layer: one sharded product step, B arriving in column panels (1-based inclusive column ranges start at col_off[p]+1);
# ready / consumed are vectors of raw CUevent handles (C_NULL entries allowed), e.g. CUDA.CuEvent(...).handle
function mul_panels!(C::CuModArray{T,2}, A::CuModArray{T,2}, B::CuModArray{T,2}, col_off::Vector{Int64};
                     ready::Vector{Ptr{Cvoid}}=Ptr{Cvoid}[], consumed::Vector{Ptr{Cvoid}}=Ptr{Cvoid}[], R::Integer=0, P::Integer=0) where {T}
    np = length(col_off) - 1
    check(ccall((:gffm_gemm_panels, libgffm), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Int32, Ptr{Int64}, Ptr{Ptr{Cvoid}}, Ptr{Ptr{Cvoid}}, UInt64, UInt64),
                C.h, A.h, B.h, np, col_off, isempty(ready) ? C_NULL : ready, isempty(consumed) ? C_NULL : consumed, R, P)); C
end
gemm_block!(C, cr, cc, A, ar, ac, B, br, bc, m, n, k; R=0, P=0, mode=0, algo=0) =
    check(ccall((:gffm_gemm_block, libgffm), Int32, (Ptr{Cvoid}, Int64, Int64, Ptr{Cvoid}, Int64, Int64, Ptr{Cvoid}, Int64, Int64, Int64, Int64, Int64, UInt64, UInt64, Int32, Int32),
                C.h, cr - 1, cc - 1, A.h, ar - 1, ac - 1, B.h, br - 1, bc - 1, m, n, k, R, P, mode, algo))
*(A::CuModArray{T,2}, B::CuModArray{T,2}) where {T} = mul!(_create(T, 2, rows(A), cols(B), A.N; ctx=A.ctx), A, B)      # kernel_ops/mul_ops.jl:54-58
*(A::CuModArray{T,2}, x::CuModArray{T,1}) where {T} = mul!(_create(T, 1, rows(A), 1, A.N; ctx=A.ctx), A, x)
mat_mul_gpu_type(A, B, mod_N::Integer=0) = mul!(_create(eltype(A), 2, rows(A), cols(B), mod_N == 0 ? A.N : mod_N; ctx=A.ctx), A, B; N=(mod_N == 0 ? A.N : mod_N))
mat_mul_type_inplace!(C, A, B, mod_N::Integer=0) = mul!(C, A, B; N=(mod_N == 0 ? C.N : mod_N))
function ^(A::CuModArray{T,2}, n::Integer) where {T}      # reference CuModMatrix.jl:307-329
    rows(A) == cols(A) || throw(CuModMatrixNotSquareException("power of a non-square matrix"))
    n < 0 && return inverse(A)^(-n)
    result = eye(T, rows(A), A.N); base = copy(A)
    while n > 0
        (n & 1) == 1 && (result = result * base)
        n >>= 1
        n > 0 && (base = base * base)
    end
    result
end

# ---- elimination (reference rref_lu_pluq/*.jl, triangular/*.jl, CuModMatrix.jl:335-502) -----------------------------------
_tuples(buf, n) = [(Int(buf[2k-1]), Int(buf[2k])) for k in 1:n]
function pluq_gpu_kernel(A::CuModArray{T,2}; debug::Bool=false, col_pivot_mode::Integer=0) where {T}
    U = Ref{Ptr{Cvoid}}(C_NULL); L = Ref{Ptr{Cvoid}}(C_NULL)
    cap = max(rows(A), cols(A), 1)
    pr = Base.zeros(Int64, 2cap); pc = Base.zeros(Int64, 2cap)
    npr = Ref{Int64}(0); npc = Ref{Int64}(0); rk = Ref{Int64}(0)
    check(ccall((:gffm_pluq, libgffm), Int32, (Ptr{Cvoid}, Ref{Ptr{Cvoid}}, Ref{Ptr{Cvoid}}, Ptr{Int64}, Ref{Int64}, Ptr{Int64}, Ref{Int64}, Ref{Int64}, Int32),
                A.h, U, L, pr, npr, pc, npc, rk, col_pivot_mode))
    (CuModArray{T,2}(U[], A.N, A.ctx), CuModArray{T,2}(L[], A.N, A.ctx), _tuples(pr, npr[]), _tuples(pc, npc[]))
end
const pluq = pluq_gpu_kernel
_setup_PLUQ(A; debug::Bool=false) = pluq_gpu_kernel(A; debug=debug)        # reference CuModMatrix.jl:335-338
function lu(A::CuModArray{T,2}) where {T}                                # intended lu_gpu_type, test/Experiments/rref_gpu_type.jl:60-103
    U = Ref{Ptr{Cvoid}}(C_NULL); L = Ref{Ptr{Cvoid}}(C_NULL); cap = max(min(rows(A), cols(A)), 1)
    pr = Base.zeros(Int64, 2cap); piv = Base.zeros(Int64, cap); npr = Ref{Int64}(0); rk = Ref{Int64}(0)
    check(ccall((:gffm_lu, libgffm), Int32, (Ptr{Cvoid}, Ref{Ptr{Cvoid}}, Ref{Ptr{Cvoid}}, Ptr{Int64}, Ref{Int64}, Ptr{Int64}, Ref{Int64}), A.h, U, L, pr, npr, piv, rk))
    (CuModArray{T,2}(U[], A.N, A.ctx), CuModArray{T,2}(L[], A.N, A.ctx), _tuples(pr, npr[]))
end
function rref(A::CuModArray{T,2}) where {T}                              # intended rref_gpu_type, rref_gpu_type.jl:8-51
    R = Ref{Ptr{Cvoid}}(C_NULL); cap = max(min(rows(A), cols(A)), 1); piv = Base.zeros(Int64, cap); rk = Ref{Int64}(0)
    check(ccall((:gffm_rref, libgffm), Int32, (Ptr{Cvoid}, Ref{Ptr{Cvoid}}, Ptr{Int64}, Ref{Int64}), A.h, R, piv, rk))
    CuModArray{T,2}(R[], A.N, A.ctx)
end
rank(A::CuModArray) = (r = Ref{Int64}(0); check(ccall((:gffm_rank, libgffm), Int32, (Ptr{Cvoid}, Ref{Int64}), A.h, r)); Int(r[]))
function is_invertible_with_inverse(A::CuModArray{T,2}; debug::Bool=false) where {T}     # reference CuModMatrix.jl:356-422
    out = Ref{Ptr{Cvoid}}(C_NULL); ok = Ref{Int32}(0)
    check(ccall((:gffm_inverse, libgffm), Int32, (Ptr{Cvoid}, Ref{Ptr{Cvoid}}, Ref{Int32}), A.h, out, ok))
    ok[] == 0 ? (false, nothing) : (true, CuModArray{T,2}(out[], A.N, A.ctx))
end
is_invertible(A::CuModArray) = rows(A) == cols(A) && rank(A) == rows(A)                  # :460-465
function inverse(A::CuModArray; debug::Bool=false)                                       # :480-502
    ok, inv = is_invertible_with_inverse(A)
    ok || throw(MatrixNotInvertibleException("matrix is not invertible"))
    inv
end
function _triinv(A::CuModArray{T,2}, upper::Bool) where {T}
    out = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:gffm_triinv, libgffm), Int32, (Ptr{Cvoid}, Int32, Ref{Ptr{Cvoid}}), A.h, upper ? 1 : 0, out))
    CuModArray{T,2}(out[], A.N, A.ctx)
end
upper_triangular_inverse_no_copy(A; debug::Bool=false) = _triinv(A, true)    # triangular_inverse_no_copy.jl:197-228
lower_triangular_inverse_no_copy(A; debug::Bool=false) = _triinv(A, false)   # :450-478
backward_sub_gpu_type_32(A) = _triinv(A, true)                               # substitution_inplace.jl:51-56
forward_sub_gpu_type_32(A) = _triinv(A, false)                               # :37-43
function _perm!(A, P::Vector{Tuple{Int,Int}}, on_cols::Bool, inv::Bool)      # permutations.jl:11-27,73-89
    flat = Int64[x for t in P for x in t]
    check(ccall((:gffm_apply_perm, libgffm), Int32, (Ptr{Cvoid}, Ptr{Int64}, Int64, Int32, Int32), A.h, flat, length(P), on_cols ? 1 : 0, inv ? 1 : 0)); A
end
apply_col_perm!(P, A) = _perm!(A, P, true, false); apply_col_inv_perm!(P, A) = _perm!(A, P, true, true)
apply_row_perm!(P, A) = _perm!(A, P, false, false); apply_row_inv_perm!(P, A) = _perm!(A, P, false, true)
function perm_array_to_matrix(perm::Vector, N::Integer, new_size::Tuple{Int,Int}; perm_stack::Bool=false)   # permutations.jl:141-157
    n = length(perm)
    if perm_stack
        P = Matrix{Int}(I, n, n); for (i, j) in perm; P[i, :], P[j, :] = P[j, :], P[i, :]; end
    else
        P = Base.zeros(Int, n, n); for i in 1:n; P[perm[i], i] = 1; end
    end
    CuModMatrix(P, N; new_size=new_size)
end
function mod_inv(p::Integer, P::Integer)                                      # pluq_kernels.jl:11-31 (batched device kernel)
    i = UInt64[mod(p, P)]; o = UInt64[0]
    check(ccall((:gffm_modinv_batch, libgffm), Int32, (Ptr{Cvoid}, Ptr{UInt64}, Ptr{UInt64}, Int64, UInt64), default_context().h, i, o, 1, P)); Int(o[1])
end

# ---- Karatsuba two-limb matrices (reference src/KaratsubaMatrix/*.jl) --------------------------------------------------
mutable struct KaratsubaArray{T,D}
    data1::CuModArray{T,D}; data2::CuModArray{T,D}; plan::Union{Nothing,CuModArray{T,D}}
    N1::Int; N2::Int; M::Int
end
const KaratsubaMatrix{T} = KaratsubaArray{T,2}
const KaratsubaVector{T} = KaratsubaArray{T,1}
KaratsubaMatrix(d1::CuModArray{T,2}, d2::CuModArray{T,2}, N1, N2, M=N1 * N2) where {T} = KaratsubaArray{T,2}(d1, d2, nothing, N1, N2, M)
KaratsubaVector(d1::CuModArray{T,1}, d2::CuModArray{T,1}, N1, N2, M=N1 * N2) where {T} = KaratsubaArray{T,1}(d1, d2, nothing, N1, N2, M)
function KaratsubaMatrix(::Type{T}, A::AbstractMatrix, N1, N2, M=N1 * N2) where {T}       # KaratsubaMatrix.jl:372-397
    Am = mod.(A, M); KaratsubaMatrix(CuModMatrix(mod.(Am, N1), N1; elem_type=T), CuModMatrix(div.(Am, N1), N1; elem_type=T), N1, N2, M)
end
MatToKMat(::Type{T}, A::AbstractArray, M::Integer) where {T} = KaratsubaMatrix(T, A, M, M, M)                  # :367-370
MatToKMat(A::AbstractArray, M::Integer) = MatToKMat(eltype(A), A, M)                                          # :358-360
MatToKMat(::Type{T}, A::AbstractArray, N1::Integer, N2::Integer, M::Integer=N1 * N2) where {T} = KaratsubaMatrix(T, A, N1, N2, M)
KMatToMat(::Type, K::KaratsubaArray) = Array(K)                                                               # :352-356
Base.size(K::KaratsubaArray) = size(K.data1)                                                                  # :302
Base.getindex(K::KaratsubaArray, i::Int, j::Int) = Int(K.data1[i, j]) + K.N1 * Int(K.data2[i, j])             # :304
Base.getindex(K::KaratsubaArray, i::Int) = Int(K.data1[i]) + K.N1 * Int(K.data2[i])                           # :305
Base.setindex!(K::KaratsubaArray, v, i::Int, j::Int) = (K.data1[i, j] = rem(v, K.N1); K.data2[i, j] = div(v, K.N1))   # :307-310
Base.setindex!(K::KaratsubaArray, v, i::Int) = (K.data1[i] = rem(v, K.N1); K.data2[i] = div(v, K.N1))         # :312-316
Base.copy!(B::KaratsubaArray, A::KaratsubaArray) = (copy!(B.data1, A.data1); copy!(B.data2, A.data2); B)      # :338-345
zero!(K::KaratsubaArray) = (zero!(K.data1); zero!(K.data2); K)                                                # :347-350
Karatsubacopy(A::KaratsubaArray{T,D}) where {T,D} = KaratsubaArray{T,D}(copy(A.data1), copy(A.data2), nothing, A.N1, A.N2, A.N1 * A.N2)   # :738-744
_klike(A::KaratsubaArray{T,D}, r=rows(A.data1), c=cols(A.data1)) where {T,D} = KaratsubaZeros(T, r, c, A.N1, A.N2, A.M)
KaratsubaZeros(::Type{T}, r, c, N1, N2, M=N1 * N2, use_gpu::Bool=true) where {T} = KaratsubaMatrix(zeros(T, r, c, N1), zeros(T, r, c, N1), N1, N2, M)   # :404-420
initialize_plan!(K::KaratsubaArray) = K        # :422-424 -- the limb add is fused into the GEMM prologue, no plan buffer needed
Array(K::KaratsubaArray) = Int.(Array(K.data1)) .+ K.N1 .* Int.(Array(K.data2))            # :318-336
function KMatMul!(C::KaratsubaArray, A::KaratsubaArray, B::KaratsubaArray)                 # :133-204 (and KMatMul_gemv! :238-300)
    (A.M == B.M == C.M) || throw(CuModArrayModulusMismatchException("Karatsuba operands have different moduli"))
    check(ccall((:gffm_kmat_mul, libgffm), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, UInt64, UInt64),
                C.data1.h, C.data2.h, A.data1.h, A.data2.h, B.data1.h, B.data2.h, A.N1, A.N2)); C
end
const KMatMul_gemv! = KMatMul!
_kew(op, C, A, B, s) = (check(ccall((:gffm_kmat_ewise, libgffm), Int32, (Int32, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Int64, UInt64, UInt64),
    op, C.data1.h, C.data2.h, A.data1.h, A.data2.h, B === nothing ? C_NULL : B.data1.h, B === nothing ? C_NULL : B.data2.h, Int64(s), A.N1, A.N2)); C)
add!(C::KaratsubaArray, A::KaratsubaArray, B::KaratsubaArray) = _kew(1, C, A, B, 0)       # :505-536
sub!(C::KaratsubaArray, A::KaratsubaArray, B::KaratsubaArray) = _kew(2, C, A, B, 0)       # :583-629
scalar_multiply!(C::KaratsubaArray, A::KaratsubaArray, s::Integer) = _kew(7, C, A, nothing, s)   # :631-666
negate!(C::KaratsubaArray, A::KaratsubaArray) = _kew(6, C, A, nothing, 0)                 # :691-731
+(A::KaratsubaArray, B::KaratsubaArray) = add!(_klike(A), A, B)                           # :428-438
-(A::KaratsubaArray, B::KaratsubaArray) = sub!(_klike(A), A, B)                           # :440-449
*(a::Number, A::KaratsubaArray) = scalar_multiply!(_klike(A), A, a)                       # :451-461
*(A::KaratsubaArray, a::Number) = a * A                                                   # :463-465
*(A::KaratsubaArray, B::KaratsubaArray) = KMatMul!(_klike(A, rows(A.data1), cols(B.data1)), A, B)   # :467-503 (unfinished in the reference)

# ---- Hensel lifting of an inverse (reference src/CuModMatrix/triangular/hensel.jl:13-21; not loaded there, needs Nemo) -----
# A, T carry the modulus N^precision; returns the lifted T (device matrix; the reference wraps Array(T) in a Nemo residue ring)
function hensel_pseudoinverse(N::Integer, precision::Integer, A::CuModArray{E,2}, T::CuModArray{E,2}) where {E}
    (A.N == N^precision == T.N) || throw(CuModArrayModulusMismatchException("A and T must carry the modulus N^precision"))
    T = copy(T); W = _like(A); V = _like(A)
    i = 1
    while i < precision
        mul!(W, A, T); mul!(V, T, W)          # T*(A*T)
        mul!(W, T, 2); sub!(T, W, V)          # T = 2T - T*(A*T)
        i *= 2
    end
    T
end
function hensel_pseudoinverse!(steps::Integer, A::KaratsubaArray{E,2}, T::KaratsubaArray{E,2}) where {E}   # two-limb moduli up to 2^52
    W = KaratsubaZeros(E, rows(A.data1), cols(A.data1), A.N1, A.N2); V = KaratsubaZeros(E, rows(A.data1), cols(A.data1), A.N1, A.N2)
    for _ in 1:steps
        KMatMul!(W, A, T); KMatMul!(V, T, W); scalar_multiply!(W, T, 2); sub!(T, W, V)
    end
    T
end

# ---- multi-GPU layer of the C ABI (new; include/gffm.h gffm_mg_*): one rank per GPU (Distributed.jl worker or thread), products on row
# blocks of A and C, B on `root`; the 128-byte id travels by any channel (e.g. remotecall / MPI.bcast) -------------------------------
mutable struct MultiGpu
    h::Ptr{Cvoid}; ctx::Context; rank::Int; nranks::Int
end
# transports of mg_set_transport! (enum in include/gffm.h) and the root value of an already distributed B
const MG_AUTO = 0; const MG_NCCL_BCAST = 1; const MG_NCCL_PLANES = 2; const MG_P2P_PLANES = 3; const MG_P2P_PUSH = 4; const MG_P2P_RAW = 5
const MG_DISTRIBUTED = -1
mg_unique_id() = (id = Base.zeros(UInt8, 128); check(ccall((:gffm_mg_unique_id, libgffm), Int32, (Ptr{UInt8},), id)); id)
function MultiGpu(id::Vector{UInt8}, nranks::Integer, rank::Integer; ctx::Context=default_context())      # rank is 0-based like NCCL's
    r = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:gffm_mg_create, libgffm), Int32, (Ptr{Cvoid}, Ptr{UInt8}, Int32, Int32, Ref{Ptr{Cvoid}}), ctx.h, id, nranks, rank, r))
    MultiGpu(r[], ctx, rank, nranks)
end
Base.close(m::MultiGpu) = (m.h == C_NULL || ccall((:gffm_mg_destroy, libgffm), Int32, (Ptr{Cvoid},), m.h); m.h = C_NULL; nothing)
function mg_info(m::MultiGpu)
    a = Ref{Int32}(0); b = Ref{Int32}(0); t = Ref{Int32}(0); p = Ref{Int32}(0)
    check(ccall((:gffm_mg_info, libgffm), Int32, (Ptr{Cvoid}, Ref{Int32}, Ref{Int32}, Ref{Int32}, Ref{Int32}), m.h, a, b, t, p))
    (rank=Int(a[]), nranks=Int(b[]), transport=Int(t[]), peer_memory=p[] != 0)
end
mg_set_transport!(m::MultiGpu, t::Integer) = check(ccall((:gffm_mg_set_transport, libgffm), Int32, (Ptr{Cvoid}, Int32), m.h, t))
mg_barrier(m::MultiGpu) = check(ccall((:gffm_mg_barrier, libgffm), Int32, (Ptr{Cvoid},), m.h))
mg_owner_ranges_root_free(n::Integer, nranks::Integer, root::Integer) = (off = Base.zeros(Int64, nranks + 1); check(ccall((:gffm_mg_owner_ranges_root_free, libgffm), Int32, (Int64, Int32, Int32, Ptr{Int64}), n, nranks, root, off)); off)
mg_owner_ranges(n::Integer, nranks::Integer) = (off = Base.zeros(Int64, nranks + 1); check(ccall((:gffm_mg_owner_ranges, libgffm), Int32, (Int64, Int32, Ptr{Int64}), n, nranks, off)); off)
# mul!(C, A, B) on row blocks: C, A = this rank's row blocks, B = the matrix on root (a same-shape matrix elsewhere)
function mg_mul!(m::MultiGpu, C::CuModArray{T,2}, A::CuModArray{T,2}, B::CuModArray{T,2}; root::Integer=0, b_ready::Ptr{Cvoid}=C_NULL, R::Integer=0, P::Integer=0) where {T}
    check(ccall((:gffm_mg_gemm, libgffm), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Int32, Ptr{Cvoid}, UInt64, UInt64), m.h, C.h, A.h, B.h, root, b_ready, R, P)); C
end
function mg_KMatMul!(m::MultiGpu, C::KaratsubaArray, A::KaratsubaArray, B::KaratsubaArray; root::Integer=0, b_ready::Ptr{Cvoid}=C_NULL)
    check(ccall((:gffm_mg_kmat_mul, libgffm), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, UInt64, UInt64, Int32, Ptr{Cvoid}),
                m.h, C.data1.h, C.data2.h, A.data1.h, A.data2.h, B.data1.h, B.data2.h, A.N1, A.N2, root, b_ready)); C
end
function mg_mul!(m::MultiGpu, z::CuModArray{T,1}, A::CuModArray{T,2}, x::CuModArray{T,1}; root::Integer=0, R::Integer=0, P::Integer=0) where {T}
    check(ccall((:gffm_mg_gemv, libgffm), Int32, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Int32, UInt64, UInt64), m.h, z.h, A.h, x.h, root, R, P)); z
end

# Exactly the reference's export list (src/GPUFiniteFieldMatrices.jl:36-60).  Like there, `zeros`, `rand`, `add!`, `sub!`, `zero!`, `rank`
# (and everything this build adds: rref, lu, hensel_pseudoinverse, mul_host!, MultiGpu, mg_*, the Karatsuba helpers, the exception types)
# stay unexported and are reached as GPUFiniteFieldMatricesB200.name -- they would clash with Base / LinearAlgebra / AbstractAlgebra.
export CuModArray, CuModMatrix, CuModVector
export inverse
export KaratsubaArray, KaratsubaMatrix, KaratsubaVector
export eye
export change_modulus, change_modulus_no_alloc!
export elementwise_multiply!, negate!
export scalar_add!, scalar_sub!, rmul!, lmul!
export mod_elements!, fill!
export mat_mul_gpu_type, mat_mul_type_inplace!
export perm_array_to_matrix
export is_invertible, inverse, is_invertible_with_inverse
export apply_col_perm!, apply_row_perm!
export mod_inv
export pluq_gpu_kernel
export upper_triangular_inverse_no_copy, lower_triangular_inverse_no_copy
export forward_sub_gpu_type_32, backward_sub_gpu_type_32

end # module
