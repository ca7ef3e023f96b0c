--  BZip2.Encoding - replacement BODY that routes Encode to the b2gpu CUDA library.
--
--  The SPEC (zip_lib/bzip2-encoding.ads of Zip-Ada v.62) is kept byte for byte: same generic
--  formals (Read_Byte, More_Bytes, Write_Byte), same Encode (option, size_hint).  This body only
--  drains the input callbacks into a buffer, calls the C ABI of include/b2gpu.h, and replays the
--  result through Write_Byte in order.  Exceptions raised inside the callbacks (User_abort,
--  Compression_inefficient, ...) propagate through Encode; the handle and buffers are released in
--  an exception handler, mirroring the clean-up of the original body
--  (bzip2-encoding.adb:1128-1133, :1351-1357, :1377-1381).
--
--  NOTE: there is no Ada compiler in the build image of this repository, so this file has not
--  been compiled.  It is deliberately small; the C++ mirror host/bzip2_encoding.hpp has the same
--  logic and is what the tests drive.  Link with -lb2gpu.

with Ada.Unchecked_Deallocation;
with Interfaces.C;
with System;

package body BZip2.Encoding is

  use Interfaces, Interfaces.C;

  subtype Handle is System.Address;

  function b2_create (level, device : int; enc : access Handle) return int;
  pragma Import (C, b2_create, "b2_create");

  procedure b2_destroy (enc : Handle);
  pragma Import (C, b2_destroy, "b2_destroy");

  function b2_bound (n : Unsigned_64) return Unsigned_64;
  pragma Import (C, b2_bound, "b2_bound");

  function b2_encode_stream
    (enc       : Handle;
     input     : System.Address;
     n         : Unsigned_64;
     size_hint : Integer_64;
     output    : System.Address;
     out_cap   : Unsigned_64;
     out_len   : access Unsigned_64) return int;
  pragma Import (C, b2_encode_stream, "b2_encode_stream");

  type Byte_Array is array (Unsigned_64 range <>) of aliased Byte;
  type Byte_Array_Access is access Byte_Array;
  procedure Free is new Ada.Unchecked_Deallocation (Byte_Array, Byte_Array_Access);

  b2gpu_error : exception;

  procedure Encode
    (option    : Compression_Option := block_900k;
     size_hint : Stream_Size_Type   := unknown_size)
  is
    level : constant int :=
      (case option is
         when block_100k => 1,
         when block_400k => 4,
         when block_900k => 9);
    enc     : aliased Handle := System.Null_Address;
    inp     : Byte_Array_Access := new Byte_Array (1 .. 1_048_576);
    outp    : Byte_Array_Access := null;
    n       : Unsigned_64 := 0;
    out_len : aliased Unsigned_64 := 0;
    grown   : Byte_Array_Access;
  begin
    --  1) Drain the input callbacks (they update the Zip CRC-32 and the feedback,
    --     zip-compress-bzip2_e.adb:70-106, exactly as before).
    while More_Bytes loop
      if n = inp'Last then
        grown := new Byte_Array (1 .. inp'Last * 2);
        grown (1 .. n) := inp (1 .. n);
        Free (inp);
        inp := grown;
      end if;
      n := n + 1;
      inp (n) := Read_Byte;
    end loop;
    --  2) One call to the device encoder (whole stream: header, blocks, footer).
    if b2_create (level, 0, enc'Access) /= 0 then
      raise b2gpu_error with "b2_create failed (no CUDA device?) - there is no CPU fallback";
    end if;
    outp := new Byte_Array (1 .. b2_bound (n) + 1024 * (n / 40_000 + 16));
    if b2_encode_stream
         (enc, inp (1)'Address, n, Integer_64 (size_hint),
          outp (1)'Address, outp'Length, out_len'Access) /= 0
    then
      raise b2gpu_error with "b2_encode_stream failed";
    end if;
    --  3) Replay the result through Write_Byte, in order (may raise Compression_inefficient).
    for i in 1 .. out_len loop
      Write_Byte (outp (i));
    end loop;
    b2_destroy (enc);
    Free (inp);
    Free (outp);
  exception
    when others =>
      if System."/=" (enc, System.Null_Address) then
        b2_destroy (enc);
      end if;
      Free (inp);
      Free (outp);
      raise;
  end Encode;

end BZip2.Encoding;
