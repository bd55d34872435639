/*
 * procfield_shim.h -- what the reference's finest-level interface exchange needs around it to compile for the host:
 *   lduMatrixUpdateMatrixInterfaces.C          lduMatrix::initMatrixInterfaces / updateMatrixInterfaces (comms-type branches)
 *   processorFvPatchScalarField.C              processorFvPatchField<scalar>::initInterfaceMatrixUpdate / updateInterfaceMatrix
 *   lduAddressingFunctors.H                    matrixPatchOperation + matrixInterfaceFunctor (the update arithmetic)
 *   processorGAMGInterfaceField.C              the same exchange on the coarse GAMG levels (:94-248)
 *   GAMGInterfaceFunctors.H                    GAMGUpdateInterfaceMatrix over the interface's cell-sorted face lists
 * are included by path; declared here: lduInterfaceField, the pointer list, Pstream / UIPstream / UOPstream over an in-process
 * mailbox (all ranks of a decomposed case live in one process), processorFvPatch, and the class declarations of
 * coupledFvPatchField / processorFvPatchField with the members those files touch (coupledFvPatchField::updateInterfaceMatrix
 * is the ten-line wrapper of coupledFvPatchField.C:221-257 around the reference's matrixPatchOperation).  TEST INFRASTRUCTURE ONLY.
 */
#ifndef PROCFIELD_SHIM_H
#define PROCFIELD_SHIM_H
#define SHIM_REFERENCE_MATRIX_INTERFACES
#include <cstring>
#include <thrust/copy.h>
#include <map>
#include <streambuf>
#include <tuple>
#include <vector>

namespace Foam
{
typedef int label;
class lduInterfaceField;
class lduInterfaceFieldPtrsList
{
    std::vector<const lduInterfaceField *> v_;

public:
    lduInterfaceFieldPtrsList() {}
    explicit lduInterfaceFieldPtrsList(label n) : v_((size_t)n, nullptr) {}
    label size() const { return (label)v_.size(); }
    bool set(label i) const { return v_[(size_t)i] != nullptr; }
    void set(label i, const lduInterfaceField *p) { v_[(size_t)i] = p; }
    const lduInterfaceField &operator[](label i) const { return *v_[(size_t)i]; }
};
struct lduScheduleEntry {
    label patch;
    bool init;
};
struct lduSchedule {
    label size() const { return 0; }
    lduScheduleEntry operator[](label) const { return lduScheduleEntry{0, false}; }
};
} // namespace Foam

#include "foam_shim.h"

namespace Foam
{
template <class T> using Field = gpuField<T>;

// ---- Pstream over a mailbox: write() posts the bytes, read() registers where they go, waitRequest() delivers ----
struct Mail {
    static int &me()
    {
        static int r = 0;
        return r;
    }
    static std::map<std::tuple<int, int, int>, std::vector<char>> &box()
    {
        static std::map<std::tuple<int, int, int>, std::vector<char>> b;
        return b;
    }
    struct Pending {
        int from, to, tag;
        char *dst;
        std::streamsize n;
    };
    static std::vector<Pending> &pending()
    {
        static std::vector<Pending> p;
        return p;
    }
};
struct UPstream {
    enum commsTypes { blocking, scheduled, nonBlocking };
    static label warnComm;
    static commsTypes defaultCommsType;
    static bool floatTransfer, gpuDirectTransfer;
    static const char *commsTypeNames[3];
    static bool parRun() { return true; }
    static label nPollProcInterfaces;
    static label nRequests() { return (label)Mail::pending().size(); }
    // the ranks run one after the other here: request storage is kept until the harness resets the case
    static void resetRequests(label) {}
    static void waitRequests() {}
    static void waitRequest(label i)
    {
        Mail::Pending &p = Mail::pending()[(size_t)i];
        const std::vector<char> &m = Mail::box().at(std::make_tuple(p.from, p.to, p.tag));
        if ((std::streamsize)m.size() != p.n) throw std::runtime_error("mailbox size");
        std::memcpy(p.dst, m.data(), (size_t)p.n);
    }
};
typedef UPstream Pstream;
struct UIPstream;
struct UOPstream;
typedef UIPstream IPstream;
typedef UOPstream OPstream;
struct UIPstream {
    static label read(UPstream::commsTypes, int fromProc, char *buf, std::streamsize n, int tag, int)
    {
        Mail::pending().push_back(Mail::Pending{fromProc, Mail::me(), tag, buf, n});
        return (label)n;
    }
};
struct UOPstream {
    static bool write(UPstream::commsTypes, int toProc, const char *buf, std::streamsize n, int tag, int)
    {
        Mail::box()[std::make_tuple(Mail::me(), toProc, tag)].assign(buf, buf + n);
        return true;
    }
};
static FatalStream Info;

// ---- lduInterfaceField.H:60-160 ----
class lduInterfaceField
{
    bool updatedMatrix_ = false;

public:
    virtual ~lduInterfaceField() {}
    bool updatedMatrix() const { return updatedMatrix_; }
    bool &updatedMatrix() { return updatedMatrix_; }
    virtual bool ready() const { return true; }
    virtual void initInterfaceMatrixUpdate(scalargpuField &, const scalargpuField &, const scalargpuField &, const direction,
                                           const Pstream::commsTypes, const bool negate = false) const
    {
    }
    virtual void updateInterfaceMatrix(scalargpuField &, const scalargpuField &, const scalargpuField &, const direction,
                                       const Pstream::commsTypes, const bool negate = false) const = 0;
};

// ---- processorFvPatch.H / fvPatch.H: rank pair, face cells, patchInternalField, compressed transfers ----
class fvMeshStub
{
public:
    lduAddressing addr_;
    const lduAddressing &lduAddr() const { return addr_; }
};
class fvPatch
{
public:
    const fvMeshStub *mesh_;
    label index_;
    labelgpuList faceCells_;
    struct BoundaryMesh {
        const fvMeshStub *m;
        const fvMeshStub &mesh() const { return *m; }
    };
    virtual ~fvPatch() {}
    const char *type() const { return "patch"; }
    const char *name() const { return "patch"; }
    label index() const { return index_; }
    label size() const { return faceCells_.size(); }
    BoundaryMesh boundaryMesh() const { return BoundaryMesh{mesh_}; }
};
class processorFvPatch : public fvPatch
{
public:
    int myProcNo_, neighbProcNo_, tag_;
    int myProcNo() const { return myProcNo_; }
    int neighbProcNo() const { return neighbProcNo_; }
    int tag() const { return tag_; }
    int comm() const { return 0; }
    // fvPatch::patchInternalField(f, pif): pif[i] = f[faceCells[i]] (fvPatchTemplates.C)
    template <class T> void patchInternalField(const gpuList<T> &f, gpuField<T> &pif) const
    {
        pif.setSize(size());
        for (label i = 0; i < size(); i++) pif.data()[i] = f.data()[faceCells_.data()[i]];
    }
    // processorLduInterface::compressedSend / compressedReceive without compression (floatTransfer off)
    template <class T> void compressedSend(const Pstream::commsTypes ct, const gpuList<T> &f) const
    {
        UOPstream::write(ct, neighbProcNo_, reinterpret_cast<const char *>(f.data()), f.byteSize(), tag_, 0);
    }
    template <class T> void compressedReceive(const Pstream::commsTypes, gpuList<T> &f) const
    {
        const std::vector<char> &m = Mail::box().at(std::make_tuple(neighbProcNo_, myProcNo_, tag_));
        if ((label)m.size() != f.byteSize()) throw std::runtime_error("mailbox size");
        std::memcpy(f.data(), m.data(), m.size());
    }
};
} // namespace Foam

#include "lduAddressingFunctors.H" /* reference: matrixPatchOperation, matrixInterfaceFunctor */

namespace Foam
{
template <class Type> class coupledFvPatchField : public lduInterfaceField
{
    const fvPatch &patch_;

public:
    coupledFvPatchField(const fvPatch &p) : patch_(p) {}
    template <class... A> coupledFvPatchField(const fvPatch &p, const A &...) : patch_(p) {} // the other constructors only parse
    const fvPatch &patch() const { return patch_; }
    void evaluate(const Pstream::commsTypes) {}
    label size() const { return patch_.size(); }
    // coupledFvPatchField.C:221-257
    void updateInterfaceMatrix(scalargpuField &result, const scalargpuField &coeffs, const scalargpuField &pnf,
                               const bool negate) const
    {
        if (negate)
            matrixPatchOperation(patch().index(), result, patch().boundaryMesh().mesh().lduAddr(),
                                 matrixInterfaceFunctor<scalar, true>(coeffs.data(), pnf.data()));
        else
            matrixPatchOperation(patch().index(), result, patch().boundaryMesh().mesh().lduAddr(),
                                 matrixInterfaceFunctor<scalar, false>(coeffs.data(), pnf.data()));
    }
};

template <class Type> class processorFvPatchField : public coupledFvPatchField<Type> // processorFvPatchField.H:60-330
{
    const processorFvPatch &procPatch_;
    mutable label outstandingSendRequest_ = -1, outstandingRecvRequest_ = -1;
    mutable gpuField<scalar> scalargpuSendBuf_, scalargpuReceiveBuf_;
    mutable Field<scalar> scalarSendBuf_, scalarReceiveBuf_;

public:
    static int debug;
    processorFvPatchField(const processorFvPatch &p) : coupledFvPatchField<Type>(static_cast<const fvPatch &>(p)), procPatch_(p) {}
    const processorFvPatch &patch() const { return procPatch_; }
    virtual bool ready() const { return true; }
    virtual void initInterfaceMatrixUpdate(scalargpuField &result, const scalargpuField &psiInternal,
                                           const scalargpuField &coeffs, const direction cmpt,
                                           const Pstream::commsTypes commsType, const bool negate = false) const;
    virtual void updateInterfaceMatrix(scalargpuField &result, const scalargpuField &psiInternal, const scalargpuField &coeffs,
                                       const direction cmpt, const Pstream::commsTypes commsType,
                                       const bool negate = false) const;
};

struct tensor {
    scalar v_[9];
};
typedef gpuField<tensor> tensorField, tensorgpuField;

// ---- cyclic patches: cyclicFvPatchField.H:60-250 (its members are the reference's cyclicFvPatchField.C) ----
class cyclicFvPatch : public fvPatch
{
public:
    const cyclicFvPatch *nbr_ = nullptr;
    label nbrID_ = -1;
    struct Nbr {
        const cyclicFvPatch *p;
        const labelgpuList &getFaceCells() const { return p->faceCells_; }
    };
    struct Poly {
        const cyclicFvPatch *p;
        Nbr neighbPatch() const { return Nbr{p->nbr_}; }
    };
    Poly cyclicPatch() const { return Poly{this}; }
    label neighbPatchID() const { return nbrID_; }
    const cyclicFvPatch &neighbFvPatch() const { return *nbr_; }
};
template <class T> bool isA(const fvPatch &p) { return dynamic_cast<const T *>(&p) != nullptr; }
struct volMesh {
};
struct fvPatchFieldMapper {
};
struct dictionary {
};
struct Ostream {
};
template <class Type, class GeoMesh> class DimensionedField;
template <class Type, template <class> class PatchField, class GeoMesh> class GeometricField;
template <class Type> class fvPatchField
{
public:
    static void write(Ostream &) {}
};
static FatalStream FatalIOError;
#define FatalIOErrorIn(where, ios) ::Foam::FatalIOError
inline int exit(FatalStream &, int) { throw std::runtime_error("FatalError"); }
template <class A, class B, class C> struct transformBinaryFunctionSFFunctor;
class cyclicLduInterfaceField
{
public:
    virtual ~cyclicLduInterfaceField() {}
    bool doTransform() const { return false; }
    struct TensorList { // only named by the transforming branch of patchNeighbourField, which is never instantiated
        tensor operator[](label) const { return tensor(); }
    };
    TensorList forwardT() const { return TensorList(); }
    template <class F> void transformCoupleField(F &, const direction) const {} // scalars: no transformation
    template <class F> void transformCoupleField(F &) const {}
};
template <class Type> class cyclicFvPatchField : public cyclicLduInterfaceField, public coupledFvPatchField<Type>
{
    const cyclicFvPatch &cyclicPatch_;

public:
    cyclicFvPatchField(const fvPatch &, const DimensionedField<Type, volMesh> &);
    cyclicFvPatchField(const fvPatch &, const DimensionedField<Type, volMesh> &, const dictionary &);
    cyclicFvPatchField(const cyclicFvPatchField<Type> &, const fvPatch &, const DimensionedField<Type, volMesh> &,
                       const fvPatchFieldMapper &);
    cyclicFvPatchField(const cyclicFvPatchField<Type> &);
    cyclicFvPatchField(const cyclicFvPatchField<Type> &, const DimensionedField<Type, volMesh> &);
    cyclicFvPatchField(const cyclicFvPatch &p, int) : coupledFvPatchField<Type>(static_cast<const fvPatch &>(p)), cyclicPatch_(p) {} // harness
    static const char *typeName;
    const cyclicFvPatch &cyclicPatch() const { return cyclicPatch_; }
    tmp<gpuField<Type>> patchNeighbourField() const;
    const cyclicFvPatchField<Type> &neighbourPatchField() const;
    virtual void updateInterfaceMatrix(scalargpuField &result, const scalargpuField &psiInternal, const scalargpuField &coeffs,
                                       const direction cmpt, const Pstream::commsTypes commsType, const bool negate = false) const;
    virtual void updateInterfaceMatrix(gpuField<Type> &result, const gpuField<Type> &psiInternal, const scalargpuField &coeffs,
                                       const Pstream::commsTypes commsType) const;
    virtual void write(Ostream &) const;
};

// ---- coarse GAMG levels: GAMGInterface / processorGAMGInterface reduced to what the field class reads ----
typedef gpuList<scalar> scalargpuList;
template <class T> struct plusEqOp { // ops.H: x += y
    void operator()(T &x, const T &y) const { x += y; }
};
template <class T> struct minusEqOp {
    void operator()(T &x, const T &y) const { x -= y; }
};
class GAMGInterface
{
public:
    labelgpuList faceCells_, sortCells_, cellFaces_, cellFacesStart_;
    virtual ~GAMGInterface() {}
    label size() const { return faceCells_.size(); }
    const labelgpuList &sortCells() const { return sortCells_; }
    const labelgpuList &cellFaces() const { return cellFaces_; }
    const labelgpuList &cellFacesStart() const { return cellFacesStart_; }
    // GAMGInterfaceTemplates.C:45-68
    template <class T> void interfaceInternalField(const gpuList<T> &iF, gpuList<T> &result) const
    {
        result.setSize(size());
        for (label i = 0; i < size(); i++) result.data()[i] = iF.data()[faceCells_.data()[i]];
    }
};
class processorGAMGInterface : public GAMGInterface
{
public:
    int myProcNo_, neighbProcNo_, tag_;
    tensorField T_;
    int comm() const { return 0; }
    int myProcNo() const { return myProcNo_; }
    int neighbProcNo() const { return neighbProcNo_; }
    int tag() const { return tag_; }
    const tensorField &forwardT() const { return T_; }
    const tensorgpuField &getForwardT() const { return T_; }
    template <class T> void compressedSend(const Pstream::commsTypes ct, const gpuList<T> &f) const
    {
        UOPstream::write(ct, neighbProcNo_, reinterpret_cast<const char *>(f.data()), f.byteSize(), tag_, 0);
    }
    template <class T> void compressedReceive(const Pstream::commsTypes, gpuList<T> &f) const
    {
        const std::vector<char> &m = Mail::box().at(std::make_tuple(neighbProcNo_, myProcNo_, tag_));
        if ((label)m.size() != f.byteSize()) throw std::runtime_error("mailbox size");
        std::memcpy(f.data(), m.data(), m.size());
    }
};
template <class To, class From> To &refCast(From &r) { return dynamic_cast<To &>(r); }
class processorLduInterfaceField
{
public:
    virtual ~processorLduInterfaceField() {}
    virtual bool doTransform() const = 0;
    virtual int rank() const = 0;
    // processorLduInterfaceField.C:37-62: scalars and untransformed couplings pass through
    void transformCoupleField(scalargpuField &, const direction) const
    {
        if (doTransform()) throw std::runtime_error("transformed couplings are not part of the harness");
    }
};
class GAMGInterfaceField : public lduInterfaceField
{
public:
    GAMGInterfaceField(const GAMGInterface &, const lduInterfaceField &) {}
    GAMGInterfaceField(const GAMGInterface &, const bool, const int) {}
};
#define TypeName(name)                              \
    static const char *typeName_() { return name; } \
    static const char *typeName;                    \
    static int debug
#define defineTypeNameAndDebug(Type, DebugSwitch) \
    const char *Type::typeName = Type::typeName_(); \
    int Type::debug = DebugSwitch
#define addToRunTimeSelectionTable(baseType, thisType, argNames) struct thisType##argNames##Unused
} // namespace Foam
#endif
