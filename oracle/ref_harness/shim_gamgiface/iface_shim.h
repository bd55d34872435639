/*
 * iface_shim.h -- what the reference's GAMG interface sources (GAMGInterface.C, GAMGInterfaceNew.C,
 * GAMGInterfaceTemplates.C, processorGAMGInterface.C) need around them to compile for the host: the
 * lduInterface / processorLduInterface bases, word, the stream types their write()/Istream members name,
 * HashTable / DynamicList / labelPair, and the run-time selection table in miniature.  Pulled in by
 * shim_gamgaddr/GAMGAgglomeration.H when SHIM_REAL_GAMG_INTERFACE is defined.  TEST INFRASTRUCTURE ONLY.
 *
 * Message passing: processorLduInterface::send / receive go through an in-process mailbox keyed by
 * (from rank, to rank) which the harness fills with the neighbour rank's data before the reference code asks.
 */
#ifndef SHIM_IFACE_H
#define SHIM_IFACE_H
#include "containers.h"

#include <map>
#include <string>
#include <unordered_map>
#include <utility>

namespace Foam
{
class word : public std::string
{
public:
    word() {}
    word(const char *s) : std::string(s) {}
    word(const std::string &s) : std::string(s) {}
};
struct Istream {
};
struct Ostream {
    template <class T> Ostream &operator<<(const T &) { return *this; }
};
namespace token
{
enum punctuationToken { SPACE = ' ' };
}
inline label readLabel(Istream &) { throw std::runtime_error("no streams in the harness"); }
struct tensor {
    scalar v_[9];
};
typedef List<tensor> tensorField;
typedef gpuList<tensor> tensorgpuField;
typedef gpuList<scalar> scalargpuField;
template <class T> using gpuField = gpuList<T>;
template <class T> using Field = List<T>;
struct UPstream {
    static label warnComm;
};
inline label max(const labelList &f)
{
    label m = f.size() ? f[0] : 0;
    forAll(f, i) if (f[i] > m) m = f[i];
    return m;
}

// className.H / typeInfo.H
#define TypeName(name)                               \
    static const char *typeName_() { return name; } \
    static const ::Foam::word typeName;              \
    static int debug;                                \
    virtual const ::Foam::word &type() const { return typeName; }
#define defineTypeNameAndDebug(Type, DebugSwitch)         \
    const ::Foam::word Type::typeName(Type::typeName_()); \
    int Type::debug(DebugSwitch)

// runTimeSelectionTables.H / addToRunTimeSelectionTable.H in miniature: name -> constructor function,
// filled by the static objects the reference's .C files define with addToRunTimeSelectionTable
#define declareRunTimeSelectionTable(autoPtr, baseType, argNames, argList, parList)                                \
    typedef autoPtr<baseType>(*argNames##ConstructorPtr) argList;                                                  \
    class argNames##ConstructorTable : public std::map<std::string, argNames##ConstructorPtr>                    \
    {                                                                                                              \
        typedef std::map<std::string, argNames##ConstructorPtr> Map;                                               \
                                                                                                                   \
    public:                                                                                                        \
        struct iterator : Map::iterator {                                                                          \
            iterator(typename Map::iterator i) : Map::iterator(i) {}                                               \
            argNames##ConstructorPtr operator()() const { return (*this)->second; }                               \
        };                                                                                                         \
        iterator find(const word &k) { return iterator(Map::find(k)); }                                            \
        iterator end() { return iterator(Map::end()); }                                                            \
        word sortedToc()                                                                                           \
        {                                                                                                          \
            std::string t;                                                                                         \
            for (auto i = Map::begin(); i != Map::end(); ++i) t += i->first + " ";                                 \
            return word(t);                                                                                        \
        }                                                                                                          \
    };                                                                                                             \
    static argNames##ConstructorTable *argNames##ConstructorTablePtr_;                                             \
    template <class T> class add##argNames##ConstructorToTable                                                     \
    {                                                                                                              \
    public:                                                                                                        \
        static autoPtr<baseType> New argList { return autoPtr<baseType>(new T parList); }                          \
        add##argNames##ConstructorToTable(const word &lookup = T::typeName)                                        \
        {                                                                                                          \
            if (!argNames##ConstructorTablePtr_) argNames##ConstructorTablePtr_ = new argNames##ConstructorTable;  \
            (*argNames##ConstructorTablePtr_)[lookup] = New;                                                       \
        }                                                                                                          \
    }
#define defineRunTimeSelectionTable(baseType, argNames) \
    baseType::argNames##ConstructorTable *baseType::argNames##ConstructorTablePtr_ = nullptr
#define addToRunTimeSelectionTable(baseType, thisType, argNames) \
    baseType::add##argNames##ConstructorToTable<thisType> add##thisType##argNames##ConstructorTo##baseType##Table_

// ---- lduInterface.H:60-150 (the members the GAMG sources call) ----
class lduInterface
{
public:
    virtual ~lduInterface() {}
    virtual const word &type() const = 0;
    virtual const labelgpuList &faceCells() const = 0;
    virtual tmp<labelField> interfaceInternalField(const labelUList &internalData) const = 0;
    virtual void initInternalFieldTransfer(Pstream::commsTypes, const labelUList &) const {}
    virtual tmp<labelField> internalFieldTransfer(Pstream::commsTypes, const labelUList &) const = 0;
};
class lduInterfacePtrsList
{
    std::vector<const lduInterface *> v_;

public:
    lduInterfacePtrsList() {}
    explicit lduInterfacePtrsList(label n) : v_((size_t)n, nullptr) {}
    label size() const { return (label)v_.size(); }
    bool set(label i) const { return v_[(size_t)i] != nullptr; }
    void set(label i, const lduInterface *p) { v_[(size_t)i] = p; }
    const lduInterface &operator[](label i) const { return *v_[(size_t)i]; }
};
template <class To, class From> To &refCast(From &r) { return dynamic_cast<To &>(r); }

// ---- processorLduInterface.H:50-150: rank pair + typed send/receive ----
struct Mailbox {
    static std::map<std::pair<int, int>, std::vector<label>> &box()
    {
        static std::map<std::pair<int, int>, std::vector<label>> b;
        return b;
    }
};
class processorLduInterface
{
public:
    virtual ~processorLduInterface() {}
    virtual int comm() const = 0;
    virtual int myProcNo() const = 0;
    virtual int neighbProcNo() const = 0;
    virtual const tensorField &forwardT() const = 0;
    virtual int tag() const = 0;
    template <class T> void send(Pstream::commsTypes, const List<T> &f) const
    {
        Mailbox::box()[std::make_pair(myProcNo(), neighbProcNo())].assign(f.begin(), f.end());
    }
    template <class T> void receive(Pstream::commsTypes, List<T> &f) const
    {
        const std::vector<label> &m = Mailbox::box().at(std::make_pair(neighbProcNo(), myProcNo()));
        if ((label)m.size() != f.size()) throw std::runtime_error("mailbox size");
        for (label i = 0; i < f.size(); i++) f[i] = m[(size_t)i];
    }
};

// ---- labelPair.H, HashTable.H, DynamicList.H ----
class labelPair : public std::pair<label, label>
{
public:
    labelPair() {}
    labelPair(label a, label b) : std::pair<label, label>(a, b) {}
    template <class = void> struct Hash {
        size_t operator()(const labelPair &p) const
        {
            return std::hash<long long>()(((long long)p.first << 32) ^ (unsigned)p.second);
        }
    };
};
template <class T, class Key, class H> class HashTable
{
    std::unordered_map<Key, T, H> m_;

public:
    explicit HashTable(label = 0) {}
    struct const_iterator {
        typename std::unordered_map<Key, T, H>::const_iterator i_;
        const T &operator()() const { return i_->second; }
        bool operator==(const const_iterator &o) const { return i_ == o.i_; }
        bool operator!=(const const_iterator &o) const { return i_ != o.i_; }
    };
    const_iterator find(const Key &k) const { return const_iterator{m_.find(k)}; }
    const_iterator end() const { return const_iterator{m_.end()}; }
    bool insert(const Key &k, const T &v) { return m_.emplace(k, v).second; }
};
template <class T> class DynamicList
{
    std::vector<T> v_;

public:
    explicit DynamicList(label reserve = 0) { v_.reserve((size_t)reserve); }
    label size() const { return (label)v_.size(); }
    void append(const T &x) { v_.push_back(x); }
    const T *begin() const { return v_.data(); }
    const T *end() const { return v_.data() + v_.size(); }
    void clear() { v_.clear(); }
};

class GAMGInterface;
class processorGAMGInterface;
} // namespace Foam
#endif
